// Microbenchmark: issue rate of FADD vs FADD2 (packed fp32) and FMNMX on sm_100a, 1..4 warps per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
template <int MODE>
__global__ void k(float* out, long long* cyc, float e, int iters) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; p[i] = pk(a[i], a[i] + 1); }
    const unsigned long long ee = pk(e, e);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] += e;                                              // FADD
                if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ee));   // FADD2
                if (MODE == 2) a[i] = fmaxf(a[i], e + i);                              // FMNMX
                if (MODE == 3) { a[i] += e; p[i] = p[i]; asm volatile("" : "+l"(p[i])); }
                if (MODE == 4) { if (i & 1) a[i] += e; else a[i] = fmaxf(a[i], e); }   // alternate FADD / FMNMX
                if (MODE == 5) { if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ee)); else a[i] = fmaxf(a[i], e); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); s += a[i] + x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char* names[] = {"FADD", "FADD2", "FMNMX", "FADD(b)", "FADD/FMNMX alt", "FADD2/FMNMX alt"};
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int m = 0; m < 6; ++m) {
            if (m == 3) continue;
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (m) {
                    case 0: k<0><<<1, warps * 32>>>(out, cyc, 1.0f, iters); break;
                    case 1: k<1><<<1, warps * 32>>>(out, cyc, 1.0f, iters); break;
                    case 2: k<2><<<1, warps * 32>>>(out, cyc, 1.0f, iters); break;
                    case 4: k<4><<<1, warps * 32>>>(out, cyc, 1.0f, iters); break;
                    case 5: k<5><<<1, warps * 32>>>(out, cyc, 1.0f, iters); break;
                }
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            const double n = (double)iters * 32;   // warp instructions per warp
            printf("%2d warps/SM (%d per SMSP) %-16s %.2f cycles per warp-instruction per SMSP\n", warps, warps / 4, names[m], h / (n * (warps / 4)));
        }
    return 0;
}
