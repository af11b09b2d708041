// bfa_host.cu -- host-buffer entry point: chunked H2D -> device pipeline -> D2H on two streams.
//
// This is the call a CPU-side owner of the posteriors makes (core.py:902-937 with tensors on the
// host).  Chunks of utterances alternate between two slots (stream + device arena), so the copy
// of chunk i+1 overlaps the kernels and the result read-back of chunk i.  The small index arrays of
// ALL chunks (lengths, targets, offsets) go up in ONE copy per call from a pinned staging buffer, so
// the copy engine sees nothing but the chunks' posteriors back to back (five small copies per chunk
// cost it ~80 us each time).
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
    // index arrays of the whole call: pinned staging, device copy, the stream that uploads them, the event the chunks wait for
    char* meta_pin = nullptr;
    char* meta_dev = nullptr;
    size_t meta_bytes = 0;
    cudaStream_t meta_stream = nullptr;
    cudaEvent_t meta_ready = nullptr;
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

int ensure_meta(Arena& a, size_t bytes) {
    if (!a.meta_stream) HTRY(cudaStreamCreateWithFlags(&a.meta_stream, cudaStreamNonBlocking));
    if (!a.meta_ready) HTRY(cudaEventCreateWithFlags(&a.meta_ready, cudaEventDisableTiming));
    if (a.meta_bytes < bytes) {
        if (a.meta_pin) HTRY(cudaFreeHost(a.meta_pin));
        if (a.meta_dev) HTRY(cudaFree(a.meta_dev));
        a.meta_pin = a.meta_dev = nullptr; a.meta_bytes = 0;
        const size_t want = bytes + bytes / 4;
        HTRY(cudaMallocHost((void**)&a.meta_pin, want));
        HTRY(cudaMalloc((void**)&a.meta_dev, want));
        a.meta_bytes = want;
    }
    return BFA_OK;
}

void release_all(Arena& a) {
    for (Slot& s : a.slot) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        if (s.dev) cudaFree(s.dev);
        if (s.pin) cudaFreeHost(s.pin);
        s = Slot();
    }
    if (a.meta_stream) { cudaStreamSynchronize(a.meta_stream); cudaStreamDestroy(a.meta_stream); }
    if (a.meta_ready) cudaEventDestroy(a.meta_ready);
    if (a.meta_pin) cudaFreeHost(a.meta_pin);
    if (a.meta_dev) cudaFree(a.meta_dev);
    a.meta_pin = a.meta_dev = nullptr; a.meta_bytes = 0; a.meta_stream = nullptr; a.meta_ready = nullptr;
}

}  // namespace

extern "C" {

void bfa_host_release(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_arena.device >= 0) cudaSetDevice(g_arena.device);
    release_all(g_arena);
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
        if (g_arena.device >= 0) release_all(g_arena);
        g_arena.device = device;
    }
    if (chunk_utts <= 0) {
        // default: ~64 MB of posteriors per chunk, at least 256 utterances, at least 4 chunks in flight when possible
        double avg = (double)shape->total_frames * C * 4.0 / (double)B;
        long long c = (long long)(64.0 * 1024 * 1024 / (avg > 1 ? avg : 1));
        chunk_utts = (int)std::max<long long>(256, std::min<long long>(c, B));
    }
    chunk_utts = std::min(chunk_utts, B);

    // ---- index arrays of the whole call, staged once: T | targets | row offsets relative to each chunk's first row |
    //      per chunk (Bc + 1) target offsets and frame offsets relative to the chunk
    const int n_chunks = (B + chunk_utts - 1) / chunk_utts;
    const long long n_tgt_all = tgt_off[B] - tgt_off[0];
    const size_t m_T = 0;
    const size_t m_tgt = al((size_t)B * 4);
    const size_t m_rowoff = m_tgt + al((size_t)std::max<long long>(n_tgt_all, 1) * 4);
    const size_t m_tgtoff = m_rowoff + al((size_t)B * 8);
    const size_t m_foff = m_tgtoff + al((size_t)(B + n_chunks) * 8);
    const size_t m_total = m_foff + al((size_t)(B + n_chunks) * 8);
    int rc = ensure_meta(g_arena, m_total);
    if (rc) return rc;
    {
        char* mp = g_arena.meta_pin;
        memcpy(mp + m_T, T, (size_t)B * 4);
        if (n_tgt_all > 0) memcpy(mp + m_tgt, tgt + tgt_off[0], (size_t)n_tgt_all * 4);
        int64_t* h_rowoff = (int64_t*)(mp + m_rowoff);
        int64_t* h_tgtoff = (int64_t*)(mp + m_tgtoff);
        int64_t* h_foff = (int64_t*)(mp + m_foff);
        for (int u0 = 0, c = 0; u0 < B; u0 += chunk_utts, ++c) {
            const int Bc = std::min(chunk_utts, B - u0);
            long long acc = 0;
            for (int i = 0; i < Bc; ++i) { h_rowoff[u0 + i] = acc; acc += (long long)T[u0 + i] * C; }
            for (int i = 0; i <= Bc; ++i) {
                h_tgtoff[u0 + c + i] = tgt_off[u0 + i] - tgt_off[u0];
                h_foff[u0 + c + i] = frame_off[u0 + i] - frame_off[u0];
            }
        }
        HTRY(cudaMemcpyAsync(g_arena.meta_dev, mp, m_total, cudaMemcpyHostToDevice, g_arena.meta_stream));
        HTRY(cudaEventRecord(g_arena.meta_ready, g_arena.meta_stream));
    }
    const char* md = g_arena.meta_dev;
    const int64_t* h_rowoff_all = (const int64_t*)(g_arena.meta_pin + m_rowoff);

    int ci = 0;
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
        size_t ws = bfa_workspace_bytes(p, &cs);
        if (ws == 0) { rc = BFA_E_UNSUPPORTED; break; }
        // device carve-up (the index arrays live in the call-wide buffer)
        size_t o = 0;
        const size_t o_lp = o; o = al(o + (size_t)lp_elems * 4);
        const size_t o_ph = o; o = al(o + (size_t)cs.total_frames * 4);
        const size_t o_ix = o; o = al(o + (size_t)cs.total_frames * 4);
        const size_t o_dpf = o; o = al(o + (size_t)Bc * 4);
        const size_t o_st = o; o = al(o + (size_t)Bc * 4);
        const size_t o_stamps = o; o = al(o + (stamps ? (size_t)Bc * ms * sizeof(BfaStamp) : 0));
        const size_t o_conf = o; o = al(o + (conf ? (size_t)Bc * ms * 4 : 0));
        const size_t o_ns = o; o = al(o + (size_t)Bc * 4);
        const size_t o_ws = o; o = al(o + ws);
        rc = ensure(s, o, 256);
        if (rc) break;
        // the slot's previous chunk must have drained before its staging memory is rewritten
        HTRY(cudaStreamSynchronize(s.stream));
        char* d = s.dev;
        cudaStream_t st = s.stream;
        if (contiguous) {
            HTRY(cudaMemcpyAsync(d + o_lp, logp + row_off[u0], (size_t)lp_elems * 4, cudaMemcpyHostToDevice, st));
        } else {
            for (int i = 0; i < Bc; ++i)
                HTRY(cudaMemcpyAsync(d + o_lp + (size_t)h_rowoff_all[u0 + i] * 4, logp + row_off[u0 + i], (size_t)T[u0 + i] * C * 4,
                                     cudaMemcpyHostToDevice, st));
        }
        if (ci < 2) HTRY(cudaStreamWaitEvent(st, g_arena.meta_ready, 0));      // later chunks of a slot are ordered behind its first
        rc = bfa_align_batch(p, &cs, (const float*)(d + o_lp), (const int64_t*)(md + m_rowoff) + u0, (const int32_t*)(md + m_T) + u0,
                             (const int32_t*)(md + m_tgt) + (tgt_off[u0] - tgt_off[0]), (const int64_t*)(md + m_tgtoff) + u0 + ci,
                             (int32_t*)(d + o_ph), (int32_t*)(d + o_ix), (const int64_t*)(md + m_foff) + u0 + ci, (float*)(d + o_dpf),
                             (int32_t*)(d + o_st), stamps ? (BfaStamp*)(d + o_stamps) : nullptr, conf ? (float*)(d + o_conf) : nullptr,
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
    if (g_arena.meta_stream) cudaStreamSynchronize(g_arena.meta_stream);
    for (Slot& s : g_arena.slot)
        if (s.stream) {
            cudaError_t e = cudaStreamSynchronize(s.stream);
            if (e != cudaSuccess && rc == BFA_OK) { snprintf(g_err, sizeof(g_err), "%s", cudaGetErrorString(e)); rc = BFA_E_CUDA; }
        }
    return rc;
}

const char* bfa_host_last_error(void) { return g_err; }

}  // extern "C"
