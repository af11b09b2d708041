// Microbenchmark: issue rates of FMNMX / SHF / LOP3 / FADD / FADD2 and mixes on sm_100a (4 warps per SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, float e, int iters) {
    float a[8]; uint32_t w[8]; unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; w[i] = threadIdx.x * 77 + i; asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 1)); }
    unsigned long long ee; asm("mov.b64 %0, {%1,%2};" : "=l"(ee) : "f"(e), "f"(e));
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
                if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ee));
                if (MODE == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
                if (MODE == 3) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(w[i]) : "r"(w[(i + 1) & 7]));
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(w[(i + 1) & 7]), "r"(w[(i + 2) & 7]));
                if (MODE == 5) { if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e)); else asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e)); }
                if (MODE == 6) { if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ee)); else asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e)); }
                if (MODE == 7) { if (i & 1) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(w[i]) : "r"(w[(i + 1) & 7])); else asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e)); }
                if (MODE == 8) { if ((i % 3) == 0) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ee)); else if ((i % 3) == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e)); else asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(w[i]) : "r"(w[(i + 1) & 7])); }
                if (MODE == 9) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += a[i] + x + y + w[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
#define RUN(M) case M: k<M><<<1, warps * 32>>>(out, cyc, 1.0f, iters); break;
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char* names[] = {"FADD", "FADD2", "FMNMX", "SHF", "LOP3", "FADD|FMNMX", "FADD2|FMNMX", "SHF|FMNMX", "FADD2|FMNMX|SHF", "MUFU.EX2"};
    for (int warps = 4; warps <= 16; warps *= 4)
        for (int m = 0; m < 10; ++m) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (m) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) }
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("%d warp(s) per SMSP  %-16s %.2f cycles per warp-instruction per SMSP\n", warps / 4, names[m], h / ((double)iters * 32 * (warps / 4)));
        }
    return 0;
}
