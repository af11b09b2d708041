// bfa_host.cu -- host-buffer entry point: chunked H2D -> device pipeline -> D2H on two streams.
//
// This is the call a CPU-side owner of the posteriors makes (core.py:902-937 with tensors on the
// host).  Chunks of utterances alternate between two slots (stream + device arena), so the copy
// of chunk i+1 overlaps the kernels and the result read-back of chunk i.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "../../include/bfa_b200.h"

namespace {

struct Slot {
    cudaStream_t stream = nullptr;
    char* dev = nullptr;
    size_t dev_bytes = 0;
    char* pin = nullptr;      // pinned staging for the small index arrays
    size_t pin_bytes = 0;
};

struct Arena {
    int device = -1;
    Slot slot[2];
};

std::mutex g_mu;
Arena g_arena;
char g_err[256] = "";

inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

#define HTRY(expr)                                                                                  \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            snprintf(g_err, sizeof(g_err), "%s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BFA_E_CUDA;                                                                      \
        }                                                                                           \
    } while (0)

int ensure(Slot& s, size_t dev_bytes, size_t pin_bytes) {
    if (!s.stream) HTRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (s.dev_bytes < dev_bytes) {
        HTRY(cudaStreamSynchronize(s.stream));
        if (s.dev) HTRY(cudaFree(s.dev));
        s.dev = nullptr; s.dev_bytes = 0;
        size_t want = dev_bytes + dev_bytes / 8;
        HTRY(cudaMalloc((void**)&s.dev, want));
        s.dev_bytes = want;
    }
    if (s.pin_bytes < pin_bytes) {
        HTRY(cudaStreamSynchronize(s.stream));
        if (s.pin) HTRY(cudaFreeHost(s.pin));
        s.pin = nullptr; s.pin_bytes = 0;
        HTRY(cudaMallocHost((void**)&s.pin, pin_bytes * 2));
        s.pin_bytes = pin_bytes * 2;
    }
    return BFA_OK;
}

}  // namespace

extern "C" {

void bfa_host_release(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_arena.device >= 0) cudaSetDevice(g_arena.device);
    for (Slot& s : g_arena.slot) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        if (s.dev) cudaFree(s.dev);
        if (s.pin) cudaFreeHost(s.pin);
        s = Slot();
    }
    g_arena.device = -1;
}

int bfa_align_batch_host(const BfaParams* p, const BfaShape* shape, const float* logp, const int64_t* row_off, const int32_t* T,
                         const int32_t* tgt, const int64_t* tgt_off, int32_t* frame_ph, int32_t* frame_idx,
                         const int64_t* frame_off, float* dp_final, int32_t* status, BfaStamp* stamps, float* conf,
                         int32_t* n_stamps, int32_t device, int32_t chunk_utts) {
    if (!p || !shape || !logp || !row_off || !T || !tgt_off || !frame_ph || !frame_idx || !frame_off || !status) return BFA_E_INVALID;
    if ((stamps == nullptr) != (n_stamps == nullptr)) return BFA_E_INVALID;
    const int B = shape->B, C = shape->C, ms = shape->max_stamps;
    if (B == 0) return BFA_OK;
    std::lock_guard<std::mutex> lk(g_mu);
    HTRY(cudaSetDevice(device));
    if (g_arena.device != device) {
        if (g_arena.device >= 0) {
            for (Slot& s : g_arena.slot) {
                if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
                if (s.dev) cudaFree(s.dev);
                if (s.pin) cudaFreeHost(s.pin);
                s = Slot();
            }
        }
        g_arena.device = device;
    }
    if (chunk_utts <= 0) {
        // default: ~64 MB of posteriors per chunk, at least 256 utterances, at least 4 chunks in flight when possible
        double avg = (double)shape->total_frames * C * 4.0 / (double)B;
        long long c = (long long)(64.0 * 1024 * 1024 / (avg > 1 ? avg : 1));
        chunk_utts = (int)std::max<long long>(256, std::min<long long>(c, B));
    }
    chunk_utts = std::min(chunk_utts, B);

    int rc = BFA_OK, ci = 0;
    for (int u0 = 0; u0 < B && rc == BFA_OK; u0 += chunk_utts, ++ci) {
        const int Bc = std::min(chunk_utts, B - u0);
        Slot& s = g_arena.slot[ci & 1];
        // chunk shape
        BfaShape cs = *shape;
        cs.B = Bc; cs.max_T = 0; cs.max_N = 0;
        cs.total_frames = frame_off[u0 + Bc] - frame_off[u0];
        long long lp_elems = 0;
        bool contiguous = true;
        for (int i = 0; i < Bc; ++i) {
            int u = u0 + i;
            cs.max_T = std::max(cs.max_T, T[u]);
            cs.max_N = std::max<int>(cs.max_N, (int)(tgt_off[u + 1] - tgt_off[u]));
            if (i + 1 < Bc && row_off[u + 1] != row_off[u] + (int64_t)T[u] * C) contiguous = false;
            lp_elems += (long long)T[u] * C;
        }
        const long long n_tgt = tgt_off[u0 + Bc] - tgt_off[u0];
        size_t ws = bfa_workspace_bytes(p, &cs);
        if (ws == 0) { rc = BFA_E_UNSUPPORTED; break; }
        // device carve-up
        size_t o = 0;
        const size_t o_lp = o; o = al(o + (size_t)lp_elems * 4);
        const size_t o_rowoff = o; o = al(o + (size_t)Bc * 8);
        const size_t o_T = o; o = al(o + (size_t)Bc * 4);
        const size_t o_tgt = o; o = al(o + (size_t)std::max<long long>(n_tgt, 1) * 4);
        const size_t o_tgtoff = o; o = al(o + (size_t)(Bc + 1) * 8);
        const size_t o_foff = o; o = al(o + (size_t)(Bc + 1) * 8);
        const size_t o_ph = o; o = al(o + (size_t)cs.total_frames * 4);
        const size_t o_ix = o; o = al(o + (size_t)cs.total_frames * 4);
        const size_t o_dpf = o; o = al(o + (size_t)Bc * 4);
        const size_t o_st = o; o = al(o + (size_t)Bc * 4);
        const size_t o_stamps = o; o = al(o + (stamps ? (size_t)Bc * ms * sizeof(BfaStamp) : 0));
        const size_t o_conf = o; o = al(o + (conf ? (size_t)Bc * ms * 4 : 0));
        const size_t o_ns = o; o = al(o + (size_t)Bc * 4);
        const size_t o_ws = o; o = al(o + ws);
        const size_t pin_need = al((size_t)Bc * 8) + 2 * al((size_t)(Bc + 1) * 8);
        rc = ensure(s, o, pin_need);
        if (rc) break;
        // the slot's previous chunk must have drained before its staging memory is rewritten
        HTRY(cudaStreamSynchronize(s.stream));
        int64_t* h_rowoff = (int64_t*)s.pin;
        int64_t* h_tgtoff = (int64_t*)(s.pin + al((size_t)Bc * 8));
        int64_t* h_foff = (int64_t*)(s.pin + al((size_t)Bc * 8) + al((size_t)(Bc + 1) * 8));
        long long acc = 0;
        for (int i = 0; i < Bc; ++i) { h_rowoff[i] = acc; acc += (long long)T[u0 + i] * C; }
        for (int i = 0; i <= Bc; ++i) { h_tgtoff[i] = tgt_off[u0 + i] - tgt_off[u0]; h_foff[i] = frame_off[u0 + i] - frame_off[u0]; }
        char* d = s.dev;
        cudaStream_t st = s.stream;
        if (contiguous) {
            HTRY(cudaMemcpyAsync(d + o_lp, logp + row_off[u0], (size_t)lp_elems * 4, cudaMemcpyHostToDevice, st));
        } else {
            for (int i = 0; i < Bc; ++i)
                HTRY(cudaMemcpyAsync(d + o_lp + (size_t)h_rowoff[i] * 4, logp + row_off[u0 + i], (size_t)T[u0 + i] * C * 4,
                                     cudaMemcpyHostToDevice, st));
        }
        HTRY(cudaMemcpyAsync(d + o_rowoff, h_rowoff, (size_t)Bc * 8, cudaMemcpyHostToDevice, st));
        HTRY(cudaMemcpyAsync(d + o_T, T + u0, (size_t)Bc * 4, cudaMemcpyHostToDevice, st));
        if (n_tgt > 0) HTRY(cudaMemcpyAsync(d + o_tgt, tgt + tgt_off[u0], (size_t)n_tgt * 4, cudaMemcpyHostToDevice, st));
        HTRY(cudaMemcpyAsync(d + o_tgtoff, h_tgtoff, (size_t)(Bc + 1) * 8, cudaMemcpyHostToDevice, st));
        HTRY(cudaMemcpyAsync(d + o_foff, h_foff, (size_t)(Bc + 1) * 8, cudaMemcpyHostToDevice, st));
        rc = bfa_align_batch(p, &cs, (const float*)(d + o_lp), (const int64_t*)(d + o_rowoff), (const int32_t*)(d + o_T),
                             (const int32_t*)(d + o_tgt), (const int64_t*)(d + o_tgtoff), (int32_t*)(d + o_ph), (int32_t*)(d + o_ix),
                             (const int64_t*)(d + o_foff), (float*)(d + o_dpf), (int32_t*)(d + o_st),
                             stamps ? (BfaStamp*)(d + o_stamps) : nullptr, conf ? (float*)(d + o_conf) : nullptr,
                             stamps ? (int32_t*)(d + o_ns) : nullptr, d + o_ws, ws, st);
        if (rc) break;
        const int64_t f0 = frame_off[u0];
        HTRY(cudaMemcpyAsync(frame_ph + f0, d + o_ph, (size_t)cs.total_frames * 4, cudaMemcpyDeviceToHost, st));
        HTRY(cudaMemcpyAsync(frame_idx + f0, d + o_ix, (size_t)cs.total_frames * 4, cudaMemcpyDeviceToHost, st));
        if (dp_final) HTRY(cudaMemcpyAsync(dp_final + u0, d + o_dpf, (size_t)Bc * 4, cudaMemcpyDeviceToHost, st));
        HTRY(cudaMemcpyAsync(status + u0, d + o_st, (size_t)Bc * 4, cudaMemcpyDeviceToHost, st));
        if (stamps) {
            HTRY(cudaMemcpyAsync(stamps + (size_t)u0 * ms, d + o_stamps, (size_t)Bc * ms * sizeof(BfaStamp), cudaMemcpyDeviceToHost, st));
            HTRY(cudaMemcpyAsync(n_stamps + u0, d + o_ns, (size_t)Bc * 4, cudaMemcpyDeviceToHost, st));
            if (conf) HTRY(cudaMemcpyAsync(conf + (size_t)u0 * ms, d + o_conf, (size_t)Bc * ms * 4, cudaMemcpyDeviceToHost, st));
        }
    }
    for (Slot& s : g_arena.slot)
        if (s.stream) {
            cudaError_t e = cudaStreamSynchronize(s.stream);
            if (e != cudaSuccess && rc == BFA_OK) { snprintf(g_err, sizeof(g_err), "%s", cudaGetErrorString(e)); rc = BFA_E_CUDA; }
        }
    return rc;
}

const char* bfa_host_last_error(void) { return g_err; }

}  // extern "C"
