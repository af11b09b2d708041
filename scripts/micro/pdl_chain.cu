// Microbenchmark: do back-to-back launches of a kernel that owns a whole SM (216 KB shared memory, one CTA per SM) overlap
// CTA by CTA under programmatic dependent launch?  Every CTA triggers its dependents at once, spins for a (per CTA, per step)
// different time, optionally executes griddepcontrol.wait at `wait_at` of its run, and records globaltimer stamps.
// With overlap the K steps take ~ K x mean(CTA time); without, K x max(CTA time).
// usage: pdl_chain [steps] [base_us] [jitter_pct] [wait_at_pct (-1: never)] [tail kernel between: 0/1]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned s;
    asm("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}

struct Stamp { unsigned long long t0, t1, tw; unsigned sm, pad; };

__global__ void __launch_bounds__(448, 1) work(int step, Stamp* out, int base_ns, int jitter_pct, int wait_at_pct, int* sink) {
    extern __shared__ unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = gtime();
    // deterministic per (cta, step) jitter in [-jitter, +jitter] percent
    unsigned h = (blockIdx.x * 2654435761u) ^ (step * 40503u + 12345u);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    const int j = (int)(h % (2 * jitter_pct + 1)) - jitter_pct;
    const long long dur = (long long)base_ns * (100 + j) / 100;
    const long long wait_at = wait_at_pct >= 0 ? dur * wait_at_pct / 100 : -1;
    unsigned long long tw = 0;
    bool waited = false;
    smem[threadIdx.x] = (unsigned char)step;
    while ((long long)(gtime() - t0) < dur) {
        if (!waited && wait_at >= 0 && (long long)(gtime() - t0) >= wait_at) {
            const unsigned long long a = gtime();
            asm volatile("griddepcontrol.wait;" ::: "memory");
            tw = gtime() - a;
            waited = true;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        Stamp s; s.t0 = t0; s.t1 = gtime(); s.tw = tw; s.sm = smid(); s.pad = 0;
        out[step * gridDim.x + blockIdx.x] = s;
        if (smem[1] == 77 && sink) sink[0] = 1;
    }
}

__global__ void tail(int* counter, int* sink) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (counter[0] != 0 && threadIdx.x == 0) sink[0] = counter[0];
}

template <typename... A>
cudaError_t launch_pdl(void (*k)(A...), int grid, int block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, args...);
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 10;
    const int base_us = argc > 2 ? atoi(argv[2]) : 100;
    const int jitter = argc > 3 ? atoi(argv[3]) : 8;
    const int wait_at = argc > 4 ? atoi(argv[4]) : -1;
    const int with_tail = argc > 5 ? atoi(argv[5]) : 0;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = 216 * 1024;
    cudaFuncSetAttribute(work, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    Stamp* d; cudaMalloc(&d, sizeof(Stamp) * steps * sms);
    int* sink; cudaMalloc(&sink, 64); cudaMemset(sink, 0, 64);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {      // 0: ordinary launches, 1: programmatic dependent launches
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0, st);
            for (int s = 0; s < steps; ++s) {
                if (mode == 0) work<<<sms, 448, smem, st>>>(s, d, base_us * 1000, jitter, wait_at, sink);
                else launch_pdl(work, sms, 448, smem, st, s, d, base_us * 1000, jitter, wait_at, sink);
                if (with_tail) {
                    if (mode == 0) tail<<<sms, 64, 0, st>>>(sink + 4, sink);
                    else launch_pdl(tail, sms, 64, (size_t)0, st, sink + 4, sink);
                }
            }
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<Stamp> h(steps * sms);
        cudaMemcpy(h.data(), d, sizeof(Stamp) * steps * sms, cudaMemcpyDeviceToHost);
        double mean = 0, mx = 0, gap = 0, waits = 0; int ngap = 0;
        for (int s = 0; s < steps; ++s) {
            double smx = 0;
            for (int c = 0; c < sms; ++c) {
                const Stamp& x = h[s * sms + c];
                const double dur = (double)(x.t1 - x.t0) * 1e-3;
                mean += dur; smx = std::max(smx, dur); waits += (double)x.tw * 1e-3;
            }
            mx += smx;
        }
        // gap between a CTA's end on an SM and the next step's CTA start on the same SM
        for (int s = 0; s + 1 < steps; ++s) {
            std::vector<unsigned long long> end_by_sm(256, 0);
            for (int c = 0; c < sms; ++c) end_by_sm[h[s * sms + c].sm & 255] = h[s * sms + c].t1;
            for (int c = 0; c < sms; ++c) {
                const Stamp& x = h[(s + 1) * sms + c];
                if (end_by_sm[x.sm & 255]) { gap += ((double)x.t0 - (double)end_by_sm[x.sm & 255]) * 1e-3; ++ngap; }
            }
        }
        printf("%s%s: %d steps base %d us jitter +-%d%% wait_at %d%%: total %.1f us = %.2f us/step | mean CTA %.2f us, sum of per-step max %.2f us/step, "
               "mean same-SM gap %.2f us, mean griddepcontrol.wait stall %.2f us  [%s]\n",
               mode ? "PDL" : "plain", with_tail ? "+tail" : "", steps, base_us, jitter, wait_at, ms * 1e3, ms * 1e3 / steps, mean / (steps * sms), mx / steps,
               ngap ? gap / ngap : 0.0, waits / (steps * sms), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
