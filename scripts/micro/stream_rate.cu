// Microbenchmark: how fast can 148 x 7 warps x 4 utterance streams pull [B, T, C] fp32 rows through shared memory with
// 1-D bulk async copies (the access pattern of the banded kernel's fill), without any compute?
// usage: stream_rate [rows_per_chunk] [stages]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
constexpr int C = 66, T = 600, WARPS = 7, UPW = 4;
__global__ void __launch_bounds__(WARPS * 32, 1) stream(const float* x, int B, int rows, int nst, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk_bytes = rows * C * 4;
    const int warp_bytes = nst * UPW * chunk_bytes + 64;
    unsigned char* my = smem + (size_t)warp * ((warp_bytes + 127) / 128 * 128);
    const uint32_t bar0 = smem_u32(my + nst * UPW * chunk_bytes);
    if (lane == 0) { for (int i = 0; i < nst; ++i) mbar_init(bar0 + 8 * i, UPW); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const int n_tasks = B / UPW, n_chunks = T / rows;
    float acc = 0.f;
    uint32_t phase = 0;
    for (int j = blockIdx.x + gridDim.x * warp; j < n_tasks; j += gridDim.x * WARPS) {
        const float* src = x + (size_t)(j * UPW + lane) * T * C;
        auto issue = [&](int c, int st) {
            if (lane < UPW) {
                mbar_expect_tx(bar0 + 8 * st, chunk_bytes);
                bulk_g2s(smem_u32(my + (st * UPW + lane) * chunk_bytes), src + (size_t)c * rows * C, chunk_bytes, bar0 + 8 * st, pol);
            }
        };
        for (int c = 0; c < nst && c < n_chunks; ++c) issue(c, c);
        int st = 0;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(bar0 + 8 * st, (phase >> st) & 1u); phase ^= 1u << st;
            acc += reinterpret_cast<const float*>(my + st * UPW * chunk_bytes)[lane];
            __syncwarp();
            if (c + nst < n_chunks) issue(c + nst, st);
            st = (st + 1 == nst) ? 0 : st + 1;
        }
    }
    if (acc == 12345.f) sink[0] = acc;
}
int main(int argc, char** argv) {
    const int B = 4096;
    float* x; float* sink; cudaMalloc(&x, (size_t)B * T * C * 4); cudaMalloc(&sink, 4); cudaMemset(x, 0, (size_t)B * T * C * 4);
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // third column: CTAs (fewer than 148: can a part of the SMs pull the whole HBM bandwidth?)
    const int cfg[][3] = {{8, 3, 148}, {8, 2, 148}, {4, 6, 148}, {12, 2, 148}, {20, 1, 148}, {24, 1, 148}, {8, 1, 148}, {4, 3, 148},
                          {8, 3, 111}, {8, 3, 74}, {8, 3, 37}, {8, 1, 74}, {8, 2, 74}};
    for (auto& c : cfg) {
        const int rows = c[0], nst = c[1], grid = c[2];
        const size_t smem = (size_t)WARPS * (((size_t)nst * UPW * rows * C * 4 + 64 + 127) / 128 * 128);
        if (smem > 227 * 1024 || T % rows) { printf("rows %d stages %d: skipped\n", rows, nst); continue; }
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0); stream<<<grid, WARPS * 32, smem>>>(x, B, rows, nst, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("CTAs %3d rows/chunk %2d stages %d (%5.1f KB in flight per SM): %.1f us  %.0f GB/s  %s\n", grid, rows, nst, smem / 1024.0, best * 1e3, (double)B * T * C * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
