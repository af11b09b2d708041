#!/usr/bin/env python
"""bench.py -- aligned frames/sec of the batched forced-alignment path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

A step = one pass of the whole hot path (row stats -> device planner -> Viterbi fill + back-trace ->
timestamps + confidences) over one batch of synthetic planted-peaky log-posteriors of the metric
shape B=4096, T=600, N=40, C=66 (BASELINE.json `metric`; `configs[1]` is the same shape at B=1024).
`value` = frames all ranks aligned / max-over-ranks CUDA-event time with inputs resident in HBM.
`e2e`   = the same through the host-buffer C-ABI entry (bfa_align_batch_host): pinned host inputs,
          H2D + kernels + D2H inside the timed region.
Only the cpu_baseline / --impl reference legs touch oracle/ (as the thing being timed ON THE CPU arm,
never on the product path).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "aligned frames/sec at T=600,N=40,P=66,batch=4096; 1/2/4/8 GPU vs CPU ref"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--T", type=int, default=600)
    ap.add_argument("--N", type=int, default=40)
    ap.add_argument("--C", type=int, default=66)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--unfused-conf", action="store_true", help="A/B switch: gather the confidence inputs in the stamp kernel (BFA_FLAG_UNFUSED_CONF)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="utterances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--chain", action="store_true", help="A/B: always launch the full planner chain (no BFA_FLAG_DIRECT_ONLY)")
    ap.add_argument("--no-pipeline", action="store_true", help="A/B: no BFA_FLAG_PIPELINED (launches do not overlap)")
    return ap.parse_args()


def workload_config(a, extra=None):
    cfg = {"workload": f"synthetic planted-peaky fp32 log-posteriors, batch={a.batch} utterances/GPU, T={a.T}, N={a.N}, C={a.C}, "
                       f"full decode_alignments path (boost+floor+silence-anchor check, Viterbi, assort, confidences)",
           "batch_per_gpu": a.batch, "T": a.T, "N": a.N, "C": a.C, "L": 4 * a.N + 1,
           "l2_policy": "inputs larger than L2 (649 MB/GPU per step vs 126 MB L2), no flush needed"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(B, T, N, C):
    """per frame 4*C read + 8 written; per utterance 24*N (targets + stamp records)."""
    return B * (T * (4 * C + 8) + 24 * N)


def csrc_hash():
    """sha256 (16 hex digits) over the kernel sources: ties an ncu capture to the build it was taken on."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "bournemouth-forced-aligner_b200" / "csrc").glob("*.cu*")):
        h.update(f.name.encode()); h.update(f.read_bytes())
    return h.hexdigest()[:16]


def measured_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
def cpu_arm(a, n_utts, steps, warmup, seed=1234):
    """Time the oracle port (oracle/bfa_oracle.c, all host threads) on a bounded sample of the workload."""
    import numpy as np
    import torch
    from bfa_b200 import synth
    from oracle import oracle as orc

    threads = orc.n_host_threads()
    lp, tgt, _ = synth.planted_batch(n_utts, a.T, a.N, a.C, seed=seed)
    p = orc.params(a.C - 1, 0)
    lp_np = lp.numpy(); tg = tgt.numpy().astype(np.int32).reshape(-1)
    row_off = np.arange(n_utts, dtype=np.int64) * a.T * a.C
    Ts = np.full(n_utts, a.T, np.int32); toff = np.arange(n_utts + 1, dtype=np.int64) * a.N
    run = lambda: orc.align_batch(p, lp_np, row_off, Ts, a.C, tg, toff, max_stamps=2 * a.N + 8, n_threads=threads)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return {"value": n_utts * a.T * steps / dt, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{n_utts} utterances of the workload shape x {steps} passes, oracle/bfa_oracle.c (C restatement of "
                      f"forced_alignment.py + utils._calculate_confidences), {threads} pthreads, {dt:.2f} s wall"}, dt / steps


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    threads = orc.n_host_threads()
    n = a.cpu_sample or a.batch
    base, step_s = cpu_arm(a, n, a.steps, a.warmup)
    out = {"metric": METRIC, "value": base["value"], "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference", "config": workload_config(a, {"cpu_sample_utterances": n}),
           "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))
    return 0


# ---------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        return reference_main(a)

    import numpy as np
    import torch
    import torch.distributed as dist

    import bfa_b200
    from bfa_b200 import _cabi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()

    B, T, N, Cc = a.batch, a.T, a.N, a.C
    lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=4242 + 1000 * rank, device=dev)
    au = bfa_b200.AlignmentUtils(blank_id=Cc - 1, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)
    dec = au.viterbi_decoder
    params = dec._params(True, True, True)
    if not bool((tgt == 0).any()):
        # host-side knowledge of the targets (they come from the phonemizer on the host): no silence_id anywhere, every utterance
        # is one plain stride-4 problem -> the library launches ONE kernel per batch (BFA_FLAG_DIRECT_ONLY; utterances it could not
        # finish would be flagged BFA_ST_DEFERRED, checked below), and because the posteriors are resident before the loop starts
        # the launches may overlap their ramp-up / tail (BFA_FLAG_PIPELINED)
        params.reserved |= _cabi.HINT_NO_SIL
        if not a.chain:
            params.reserved |= _cabi.FLAG_DIRECT_ONLY
            if not a.no_pipeline:
                params.reserved |= _cabi.FLAG_PIPELINED
    if a.unfused_conf:
        params.reserved |= 4
    row_off = torch.arange(B, dtype=torch.int64, device=dev) * (T * Cc)
    tgt32 = tgt.to(torch.int32).reshape(-1).contiguous()
    Ts, Ns = [T] * B, [N] * B
    gather_bufs = None
    bplan = dec.plan_batch(Ts, Ns, Cc, params=params, device=dev)   # shape metadata uploaded once, like a real serving loop
    result = [None, None]      # two result sets: the gather of one batch overlaps the alignment of the next
    pending = [None, None]
    nstep = [0]

    # ---- the gather of the packed result arrays (the path's only exchange).  Preferred: every rank PUSHES its arena into the
    #      peers' receive buffers with copy-engine peer-to-peer writes over NVLink (torch symmetric memory), on a side stream:
    #      no SM is taken from the banded kernel, which needs all of them.  Fallback: one NCCL all_gather per batch.
    pusher = None                 # sharding.PushGather once the arena size is known; False: NCCL fallback
    use_push = world > 1 and os.environ.get("BFA_GATHER", "p2p") == "p2p"

    def step():
        nonlocal gather_bufs, pusher, use_push
        i = nstep[0] & 1
        nstep[0] += 1
        if pusher is not None:
            pusher.wait(i)
        if pending[i] is not None:     # the gather that still reads this result set
            pending[i].wait()
            pending[i] = None
        r = dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=params, want_stamps=True, want_conf=True, plan=bplan, out=result[i])
        result[i] = r
        if world > 1:
            if use_push and pusher is None:   # first batch: the arena size is known now (collective call)
                try:
                    from bfa_b200.sharding import PushGather
                    pusher = PushGather(r.arena.numel(), dev)
                except Exception as e:   # noqa: BLE001
                    use_push = False
                    if rank == 0:
                        print(f"# symmetric memory unavailable ({e}); using NCCL all_gather", file=sys.stderr)
            if pusher is not None:
                pusher.push(i, r.arena)
            else:
                if gather_bufs is None:   # stamps | conf | n_stamps | status | dp_final are one allocation: one collective
                    gather_bufs = [torch.empty(world * r.arena.numel(), dtype=r.arena.dtype, device=dev) for _ in range(2)]
                pending[i] = dist.all_gather_into_tensor(gather_bufs[i], r.arena, async_op=True)
        return r

    def drain():
        for i in range(2):
            if pusher is not None:
                pusher.wait(i)
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    # Clock sampling starts here (nvidia-smi is already polling when the timed region begins).  Then the GPU is brought to its
    # sustained clocks: an idle B200 sits at 120 MHz and needs milliseconds of load to ramp up, far longer than W = 5 steps of
    # 0.15 ms -- so PREWARM_S seconds of untimed steps come first, then the contract's W warm-up steps, then (after the barrier)
    # the K timed steps, with no host-side pause in between.
    sampler = ClockSampler(local)
    sampler.start()
    PREWARM_S = 0.4
    t_pre = time.perf_counter()
    n_pre = 0
    while time.perf_counter() - t_pre < PREWARM_S:
        for _ in range(16):
            r = step()
        n_pre += 16
        drain()
        torch.cuda.synchronize()
    for _ in range(max(a.warmup, 3)):
        r = step()
    drain()
    torch.cuda.synchronize()
    if world > 1 and pusher is not None:
        # untimed check of the push gather: what landed in my receive buffer is what the peers computed
        dist.barrier()
        torch.cuda.synchronize()
        i_last = (nstep[0] - 1) & 1
        mine = result[i_last].arena.to(torch.int64).sum().reshape(1)
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, mine)
        got = pusher.recv[i_last].view(world, -1).to(torch.int64).sum(1)
        assert bool((got == sums).all()), "push gather: receive buffer does not match the peers' results"
    assert int((r.status[:B] & 7 != 0).sum()) == 0, "unexpected non-OK status on the synthetic workload"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l0 = lib.bfa_launch_count()
    step(); drain(); torch.cuda.synchronize()
    one_kernel = (lib.bfa_launch_count() - l0) == 1     # the whole step is ONE kernel: the region's events are that kernel's events
    launches0 = lib.bfa_launch_count()
    # More than one kernel per step: the dominant kernel is bracketed by CUDA events of its own (N=1: every step, N>1: every 4th
    # step - there the event records also serialise against the side-stream pushes).  One kernel per step: no events inside the
    # region, the kernel's average duration is the region's time / K.
    PROFILE_EVERY = 1 if world == 1 else 4
    lib.bfa_profile_enable(0 if one_kernel else (1 | (PROFILE_EVERY << 8)))
    lib.bfa_profile_read(None, None)
    for _ in range(max(a.warmup, 3)):      # the W warm-up steps, immediately before the timed region
        step()
    drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    th0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    host_us_per_call = (time.perf_counter() - th0) / a.steps * 1e6     # how long the host needs to enqueue one step
    drain()                    # every gather has completed inside the timed region
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    dom_ms, dom_n = C.c_float(), C.c_int32()
    lib.bfa_profile_read(C.byref(dom_ms), C.byref(dom_n))
    lib.bfa_profile_enable(0)
    time.sleep(0.15)
    clocks = sampler.stop()
    launches = lib.bfa_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * T * a.steps / (ms / 1e3)

    # ---- the same K steps once more WITHOUT the per-kernel CUDA events of the profile switch: the two event records around the
    #      banded kernel cost device time and stand between kernels that otherwise use programmatic dependent launch.  Reported
    #      beside the contract's numbers (which stay the instrumented ones), not instead of them.
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(a.steps):
        step()
    drain()
    p1.record()
    barrier()
    plain_ms = p0.elapsed_time(p1)
    if world > 1:
        t = torch.tensor([plain_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        plain_ms = float(t.item())
    plain = {"ms_per_step": plain_ms / a.steps, "value": world * B * T * a.steps / (plain_ms / 1e3), "unit": "frames/s",
             "how": "same K steps, bfa_profile_enable(0): no per-kernel events inside any step"}

    # ---- roofline of the dominant kernel (Viterbi fill + back-trace), timed by CUDA events on its stream
    assert int((r.status[:B] & 7 != 0).sum()) == 0, "an utterance of the timed region was not finished (deferred / non-OK status)"
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes(B, T, N, Cc)
    ctag = f"<{Cc},false>" if Cc == 66 else "<0,false> (run-time class count)"
    if one_kernel:
        dom_avg_ms = ms / a.steps
        kname = f"viterbi_band3_direct_kernel{ctag} (the whole step: in-kernel planning, fill, back-trace, frame labels, stamps, confidences)"
        how = (f"one kernel per step: CUDA events on the launch stream around the {a.steps} back-to-back launches of the timed region / {a.steps}"
               + ("; consecutive launches overlap ramp-up and tail SM by SM (programmatic dependent launch), see isolated_launch_ms" if params.reserved & _cabi.FLAG_PIPELINED else ""))
    else:
        dom_avg_ms = dom_ms.value / max(dom_n.value, 1)
        kname = f"viterbi_band3_direct_kernel{ctag} (first kernel of the chain: planning, fill, back-trace, stamps, confidences of every plain utterance)"
        how = f"CUDA events on the launch stream around {'every' if PROFILE_EVERY == 1 else f'every {PROFILE_EVERY}th'} launch inside the timed region, {dom_n.value} launches averaged"
    achieved = alg / (dom_avg_ms / 1e3) / 1e9 if dom_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": kname, "kernel_ms": dom_avg_ms, "kernel_share_of_step": dom_avg_ms / (ms / a.steps),
                "algorithmic_bytes_per_launch": alg, "peak_source": peak_src, "kernel_ms_how": how}
    # ---- the same kernel launched ALONE (events around every launch: nothing overlaps), and the full chain for comparison
    if rank == 0 or world > 1:
        lib.bfa_profile_enable(1); lib.bfa_profile_read(None, None)
        for _ in range(10):
            step()
        drain(); torch.cuda.synchronize()
        i_ms, i_n = C.c_float(), C.c_int32()
        lib.bfa_profile_read(C.byref(i_ms), C.byref(i_n)); lib.bfa_profile_enable(0)
        iso = i_ms.value / max(i_n.value, 1)
        roofline["isolated_launch_ms"] = iso
        roofline["isolated_launch_frac"] = (alg / (iso / 1e3) / 1e9 / peak) if iso > 0 else 0.0
    # ---- the fill phase by itself (north_star: ">= 70 % of the HBM roofline on the batched Viterbi fill"): the same kernel with
    #      the measurement switch BFA_FLAG_FILL_ONLY (rows streamed once, log-sum-exp, forward recursion, decision records
    #      written; no back-trace, no outputs), a few extra launches outside the timed region above
    if rank == 0 or world > 1:
        import copy
        p_fill = copy.copy(params)
        p_fill.reserved |= 16
        fplan = dec.plan_batch(Ts, Ns, Cc, params=p_fill, device=dev)
        scratch = [None]
        for _ in range(3):
            scratch[0] = dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=p_fill, want_stamps=True, want_conf=True, plan=fplan, out=scratch[0])
        torch.cuda.synchronize()
        lib.bfa_profile_enable(1); lib.bfa_profile_read(None, None)
        for _ in range(10):
            dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=p_fill, want_stamps=True, want_conf=True, plan=fplan, out=scratch[0])
        torch.cuda.synchronize()
        f_ms, f_n = C.c_float(), C.c_int32()
        lib.bfa_profile_read(C.byref(f_ms), C.byref(f_n)); lib.bfa_profile_enable(0)
        fill_ms = f_ms.value / max(f_n.value, 1)
        fill_bytes = B * T * 4 * Cc                       # the fill reads every row once; its records stay in L2
        roofline["fill_phase"] = {"kernel_ms": fill_ms, "achieved": fill_bytes / (fill_ms / 1e3) / 1e9 if fill_ms > 0 else 0.0,
                                  "frac": (fill_bytes / (fill_ms / 1e3) / 1e9 / peak) if fill_ms > 0 else 0.0, "bytes": fill_bytes,
                                  "how": "same kernel, BFA_FLAG_FILL_ONLY (no back-trace, no outputs), 10 launches, CUDA events"}
        del scratch
    # DRAM traffic per launch comes from an ncu capture (it cannot be measured in an unprofiled run); it is only quoted when the
    # capture was taken on the kernel sources of THIS build (hash of csrc/), otherwise null
    tf = ROOT / "profiles" / "traffic_latest.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            if tj.get("csrc_sha16") == csrc_hash():
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
                roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- e2e through the host-buffer C-ABI entry (pinned host memory in, results out)
    e2e = None
    if not a.no_e2e:
        lp_h = torch.empty((B, T, Cc), dtype=torch.float32, pin_memory=True); lp_h.copy_(lp)
        tgt_h = torch.empty(B * N, dtype=torch.int32, pin_memory=True); tgt_h.copy_(tgt32)
        ms_stamps = N + 8
        out = {k: torch.empty(s, dtype=d, pin_memory=True).numpy() for k, s, d in (
            ("frame_ph", (B * T,), torch.int32), ("frame_idx", (B * T,), torch.int32), ("dp_final", (B,), torch.float32),
            ("status", (B,), torch.int32), ("stamps", (B, ms_stamps, 4), torch.int32), ("n_stamps", (B,), torch.int32),
            ("conf", (B, ms_stamps), torch.float32))}
        ro_h = np.arange(B, dtype=np.int64) * T * Cc; T_h = np.full(B, T, np.int32); to_h = np.arange(B + 1, dtype=np.int64) * N
        run = lambda: bfa_b200.align_host(params, lp_h.numpy(), ro_h, T_h, Cc, tgt_h.numpy(), to_h, max_stamps=ms_stamps,
                                          device=local, chunk_utts=512, out=out)
        for _ in range(3):
            run()
        barrier()
        ke = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ke):
            run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = lp_h.numel() * 4 + tgt_h.numel() * 4 + B * 8 + B * 4 + 2 * (B + 1) * 8
        d2h = sum(v.nbytes for v in out.values() if hasattr(v, 'nbytes') and v is not out.get('frame_off'))
        e2e = {"value": world * B * T * ke / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": ke, "ms_per_step": dt / ke * 1e3, "api": "bfa_align_batch_host (pinned host buffers, 512-utterance chunks, 2 streams)"}
        lib.bfa_host_release()

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import oracle as orc
        n = a.cpu_sample or B
        cpu_baseline, _ = cpu_arm(a, n, 8, 1)    # ~1 s wall on 16 threads = ~17 core-seconds of CPU work

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
               "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "impl": "b200",
               "config": workload_config(a, {"gather": ("none (1 GPU)" if world == 1 else "copy-engine P2P pushes over NVLink (sharding.PushGather)"
                                                        if pusher is not None else "NCCL all_gather_into_tensor"), "sharding": f"{world} rank(s) x {B} utterances, no data-path collective; "
                                                         f"when n_gpus>1 every rank pushes its packed result arrays to all peers each step (copy-engine P2P writes over NVLink on a side stream; NCCL all_gather as fallback), completed inside the timed region"}),
               "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "uninstrumented_step": plain, "host_enqueue_us_per_step": host_us_per_call, "prewarm_steps": n_pre}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
