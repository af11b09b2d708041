// Microbenchmark / feasibility check: conditional device-side tail launch (CUDA dynamic parallelism, CDP2) from the last CTA of a
// kernel, followed in the stream by a kernel launched with programmatic stream serialization.  Questions:
//   1. does the tail-launched child (and a grandchild launched by the child) run before the next kernel in the stream passes
//      griddepcontrol.wait / before an ordinary successor starts?
//   2. what does the launch cost when it happens, and what does the mere possibility cost when it does not?
//   3. can the sequence be captured into a CUDA graph and replayed?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -rdc=true -O3 -o cdp_tail cdp_tail.cu -lcudadevrt
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void grandchild(int* data) {
    if (threadIdx.x == 0 && blockIdx.x == 0) data[2] = data[1] + 1;
}
__global__ void child(int* data) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        data[1] = data[0] + 1;
        grandchild<<<1, 32, 0, cudaStreamTailLaunch>>>(data);
    }
}
// parent: every CTA spins `ns`, the last one to finish (ticket) launches the child when cond != 0
__global__ void parent(int* data, int* ticket, int cond, int ns) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = gtime();
    while ((long long)(gtime() - t0) < ns) {}
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(ticket, 1);
        if (t == (int)gridDim.x - 1) {
            *ticket = 0;
            data[0] = 100;
            __threadfence();
            if (cond) child<<<4, 64, 0, cudaStreamTailLaunch>>>(data);
        }
    }
}
// successor launched with programmatic stream serialization: must observe the grandchild's write after griddepcontrol.wait
__global__ void successor(int* data, int* result) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) { result[0] = data[0]; result[1] = data[1]; result[2] = data[2]; }
}

template <typename... A>
cudaError_t launch_pdl(void (*k)(A...), int grid, int block, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, args...);
}

int main() {
    int *data, *ticket, *result;
    cudaMalloc(&data, 64); cudaMalloc(&ticket, 64); cudaMalloc(&result, 64);
    cudaMemset(ticket, 0, 64);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int h[3];
    for (int cond = 0; cond < 2; ++cond) {
        for (int pdl = 0; pdl < 2; ++pdl) {
            cudaMemsetAsync(data, 0, 64, st);
            cudaMemsetAsync(result, 0xff, 64, st);
            parent<<<148, 128, 0, st>>>(data, ticket, cond, 20000);
            if (pdl) launch_pdl(successor, 1, 32, st, data, result);
            else successor<<<1, 32, 0, st>>>(data, result);
            cudaError_t e = cudaStreamSynchronize(st);
            cudaMemcpy(h, result, 12, cudaMemcpyDeviceToHost);
            printf("cond %d successor %s: data seen by the successor = (%d, %d, %d)  expected (100, %d, %d)  [%s]\n", cond, pdl ? "PDL" : "plain",
                   h[0], h[1], h[2], cond ? 101 : 0, cond ? 102 : 0, cudaGetErrorString(e));
        }
    }
    // cost: 200 x (parent 20 us + successor), cond 0 / 1
    for (int cond = 0; cond < 2; ++cond) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0, st);
            for (int i = 0; i < 200; ++i) {
                launch_pdl(parent, 148, 128, st, data, ticket, cond, 20000);
                launch_pdl(successor, 1, 32, st, data, result);
            }
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("cond %d: %.2f us per (parent 20 us + successor) pair  [%s]\n", cond, ms * 1e3 / 200, cudaGetErrorString(cudaGetLastError()));
    }
    // graph capture
    {
        cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        cudaMemsetAsync(data, 0, 64, st);
        parent<<<148, 128, 0, st>>>(data, ticket, 1, 20000);
        launch_pdl(successor, 1, 32, st, data, result);
        cudaError_t e2 = cudaStreamEndCapture(st, &g);
        cudaError_t e3 = g ? cudaGraphInstantiate(&ge, g, 0) : cudaErrorUnknown;
        cudaError_t e4 = cudaErrorUnknown, e5 = cudaErrorUnknown;
        h[0] = h[1] = h[2] = -1;
        if (ge) {
            cudaMemsetAsync(result, 0xff, 64, st);
            e4 = cudaGraphLaunch(ge, st);
            e5 = cudaStreamSynchronize(st);
            cudaMemcpy(h, result, 12, cudaMemcpyDeviceToHost);
        }
        printf("graph: begin %s, end %s, instantiate %s, launch %s, sync %s: successor saw (%d, %d, %d)\n", cudaGetErrorString(e), cudaGetErrorString(e2),
               cudaGetErrorString(e3), cudaGetErrorString(e4), cudaGetErrorString(e5), h[0], h[1], h[2]);
        (void)cudaGetLastError();
    }
    return 0;
}
