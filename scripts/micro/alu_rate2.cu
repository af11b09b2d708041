// Microbenchmark 2: rates of candidate instructions for decision-bit accumulation (4 warps per SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, float e, int iters) {
    float a[8]; uint32_t w[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; w[i] = threadIdx.x * 77 + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) asm volatile("set.gt.u32.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
                if (MODE == 1) asm volatile("{.reg .u32 t; add.u32 t, %0, %0; sub.u32 %0, t, %1;}" : "+r"(w[i]) : "r"(w[(i + 1) & 7]));   // acc = 2acc - r
                if (MODE == 2) asm volatile("mad.lo.u32 %0, %0, 2, %1;" : "+r"(w[i]) : "r"(w[(i + 1) & 7]));
                if (MODE == 3) asm volatile("max.f32 %0, %0, %1; max.f32 %0, %0, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));   // FMNMX3?
                if (MODE == 4) asm volatile("max.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(w[(i + 1) & 7]));
                if (MODE == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7351;" : "+r"(w[i]) : "r"(w[(i + 1) & 7]));
                if (MODE == 6) asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; selp.u32 %0, %0, %3, p;}" : "+r"(w[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]), "r"(w[(i + 1) & 7]));
                if (MODE == 7) asm volatile("{.reg .u32 r; set.gt.u32.f32 r, %1, %2; mad.lo.u32 %0, %0, 2, r;}" : "+r"(w[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));   // candidate sequence
                if (MODE == 8) asm volatile("{.reg .f32 d; sub.f32 d, %1, %2; shf.l.wrap.b32 %0, d, %0, 1;}" : "+r"(w[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));  // current sequence
                if (MODE == 9) asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; addc.cc.u32 %0, %0, %0;}" : "+r"(w[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
#define RUN(M) case M: k<M><<<1, warps * 32>>>(out, cyc, 1.0f, iters); break;
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char* names[] = {"FSET", "2acc-r (IADD3?)", "IMAD acc*2+r", "max,max (FMNMX3?)", "VIMNMX.U32", "PRMT", "FSETP+SEL", "FSET+IMAD", "FADD+SHF", "x"};
    for (int warps = 16; warps <= 16; warps *= 4)
        for (int m = 0; m < 9; ++m) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (m) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) }
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("%d warp(s) per SMSP  %-18s %.2f cycles per sequence per SMSP\n", warps / 4, names[m], h / ((double)iters * 32 * (warps / 4)));
        }
    return 0;
}
