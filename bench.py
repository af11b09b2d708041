#!/usr/bin/env python
"""bench.py -- aligned frames/sec of the batched forced-alignment path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU arm: the reference's own implementation on the host cores)
    python bench.py --corpus 1000000 --gpus N                (BASELINE config 5: a corpus sharded over the ranks, one gather)

A step = one pass of the whole hot path (planning -> Viterbi fill + back-trace -> frame labels -> timestamps + confidences)
over one batch of synthetic planted-peaky log-posteriors of the metric shape B=4096, T=600, N=40, C=66 (BASELINE.json
`metric`; `configs[1]` is the same shape at B=1024).
`value` = frames all ranks aligned / max-over-ranks CUDA-event time with inputs resident in HBM.
`e2e`   = the same through the host-buffer C-ABI entry (bfa_align_batch_host): pinned host inputs, H2D + kernels + D2H
          inside the timed region.
`variants` (N=1) = the reference's real class counts (67 / 17), a SIL-bearing copy of the metric shape (silence anchoring
          really on) and BASELINE configs 2-4, each with its own roofline fraction.
Only the cpu_baseline / --impl reference legs touch oracle/ (as the thing being timed ON THE CPU arm, never on the product path).
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
    ap.add_argument("--no-variants", action="store_true", help="skip the extra workloads (C=67/17, SIL-bearing, configs 2-4)")
    ap.add_argument("--unfused-conf", action="store_true", help="A/B switch: gather the confidence inputs in the stamp kernel (BFA_FLAG_UNFUSED_CONF)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="utterances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--chain", action="store_true", help="A/B: always launch the full planner chain (no BFA_FLAG_DIRECT_ONLY)")
    ap.add_argument("--no-pipeline", action="store_true", help="A/B: no BFA_FLAG_PIPELINED (launches do not overlap)")
    ap.add_argument("--gather", default="peer", choices=["peer", "root", "all", "nccl"],
                    help="N>1: peer = the alignment kernels write their result arrays straight into rank 0's memory (peer-mapped "
                         "stores over NVLink: the gather to one place costs no extra step); root = every rank streams its result arrays "
                         "to rank 0 with copy-engine pushes; all = every rank pushes to every peer each step (stress variant); "
                         "nccl = one all_gather per step")
    ap.add_argument("--corpus", type=int, default=0, help="BASELINE config 5: align a corpus of this many utterances sharded over the ranks")
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t_from=None, t_to=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(lines):
            sm, mx, reasons = [], [], set()
            for _, ln in lines:
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
            return sm, mx, reasons
        # samples taken while the GPU was under load: from the start of the pre-warm loop to the end of the timed region
        # (the first 40 ms after the load starts are the clock's ramp from idle: not part of the window)
        loaded = [x for x in self.lines if (t_from is None or x[0] >= t_from + 0.04) and (t_to is None or x[0] <= t_to + 0.03)]
        sm, mx, reasons = digest(loaded if loaded else self.lines)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": "pre-warm loop + warm-up + timed region"}


def algorithmic_bytes(Ts, Ns, C_):
    """per frame 4*C read + 8 written; per utterance 24*N (targets + stamp records)."""
    return int(sum(Ts)) * (4 * C_ + 8) + 24 * int(sum(Ns))


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
def cpu_port_arm(a, n_utts, steps, warmup, seed=1234):
    """Time the oracle port (oracle/bfa_oracle.c, all host threads) on a bounded sample of the workload."""
    import numpy as np
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


def python_reference_arm(a, per_core, steps, warmup, seed=4321):
    """Time the UNMODIFIED reference (forced_alignment.py + utils.py staged under oracle/_ref by oracle/stage_reference.py):
    one worker process per host core, torch threads = 1 each, `per_core` utterances of the workload shape per worker and step.
    Returns (None, None) when the staged copy is absent."""
    from oracle import stage_reference as sr
    if not sr.available():
        return None, None
    from bfa_b200 import synth
    cores = os.cpu_count() or 1
    n = cores * per_core
    lp, tgt, _ = synth.planted_batch(n, a.T, a.N, a.C, seed=seed)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    bounds = [n * i // cores for i in range(cores + 1)]
    jobs = [(lp[bounds[i]:bounds[i + 1]].clone(), tgt[bounds[i]:bounds[i + 1]].clone(), a.T, a.N, a.C - 1, True) for i in range(cores)]
    walls = []
    with ctx.Pool(cores) as pool:
        pool.map(sr._noop, range(cores))              # workers imported torch and the staged reference
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(sr._worker, jobs)
            if it >= warmup:
                walls.append(time.perf_counter() - t0)
    dt = sum(walls)
    return {"value": n * a.T * steps / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{n} utterances of the workload shape ({per_core} per core) x {steps} passes through the unmodified "
                      f"AlignmentUtils.decode_alignments + utils._calculate_confidences (oracle/_ref, staged from the reference), "
                      f"{cores} worker processes with torch.set_num_threads(1), {dt:.2f} s wall"}, dt / steps


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    # bounded samples: the Python reference aligns ~11 utterances of this shape per second and core
    ref, ref_step = python_reference_arm(a, 2, min(steps, 6), min(warmup, 1))
    n = a.cpu_sample or a.batch
    port, port_step = cpu_port_arm(a, n, min(steps, 8), min(warmup, 1))
    base, step_s = (ref, ref_step) if ref is not None else (port, port_step)
    out = {"metric": METRIC, "value": base["value"], "unit": "frames/s", "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
           "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference", "config": workload_config(a, {"cpu_sample_utterances": n}),
           "cpu_baseline": base, "cpu_port": port,
           "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": ("value = the unmodified Python reference on every host core (kind 'reference'); cpu_port = the C restatement of the same "
                    "path on all host threads (a far stricter CPU baseline)") if ref is not None else
                   "the staged reference (oracle/_ref) is absent: value = the C restatement of the path (kind 'port')"}
    print(json.dumps(out))
    return 0


# ---------------------------------------------------------------------------------------------
def time_steps(torch, step, n, sync=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    if sync:
        sync()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_variant(torch, lib, dec, dev, name, w, Cc, params, peak, steps=10):
    """Device-timed steps of one extra workload (inputs resident in HBM); returns the bench-line entry."""
    import numpy as np
    plan = dec.plan_batch(w["Ts"], w["Ns"], Cc, params=params, device=dev)
    res = [None]

    def step():
        res[0] = dec.align_batch(w["lp"], w["row_off"], w["Ts"], Cc, w["tgt"], w["Ns"], params=params, plan=plan, out=res[0])
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    l0 = lib.bfa_launch_count()
    ms = time_steps(torch, step, steps)
    launches = (lib.bfa_launch_count() - l0) / steps
    B = len(w["Ts"])
    alg = algorithmic_bytes(w["Ts"], w["Ns"], Cc)
    frames = int(sum(w["Ts"]))
    codes, counts = np.unique((res[0].status[:B] & 15).cpu().numpy(), return_counts=True)
    ic = (C.c_int32 * 4)(); lib.bfa_debug_item_counts(ic)
    out = {"name": name, "B": B, "C": Cc, "frames": frames, "input_MB": round(w["lp"].numel() * 4 / 1e6, 1), "ms_per_step": ms,
           "value": frames / (ms / 1e3), "unit": "frames/s", "launches_per_step": launches,
           "roofline": {"algorithmic_bytes": alg, "achieved": alg / (ms / 1e3) / 1e9, "frac": alg / (ms / 1e3) / 1e9 / peak, "unit": "GB/s",
                        "of": "the whole step (every kernel of the call), CUDA events around %d steps" % steps},
           "status_counts": {str(int(k)): int(v) for k, v in zip(codes, counts)},
           "items": {"exact_kernel": int(ic[0]), "window_24_40_64": [int(ic[1]), int(ic[2]), int(ic[3])]}}
    del res
    return out


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
    if a.corpus:
        return corpus_main(a, world, rank, local, dev, lib)

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
    bplan = dec.plan_batch(Ts, Ns, Cc, params=params, device=dev)   # shape metadata uploaded once, like a real serving loop

    # ---- the gather of the packed result arrays (the path's only exchange).  Default: every rank STREAMS its arena of every
    #      step to rank 0 (copy-engine peer-to-peer writes over NVLink on a side stream, torch symmetric memory): when the last
    #      step is done rank 0 holds every rank's timestamp arrays -- north_star's final gather to one place, overlapped with
    #      the alignment, no SM taken from the kernel that needs all of them, 1/(N-1) of the traffic of an all-gather.
    #      --gather all: every rank pushes to every peer each step (round 1's design, kept as a stress variant);
    #      --gather nccl: one NCCL all_gather per step.
    n_slots = max(a.steps, 2) if (a.gather in ("root", "peer") and world > 1) else 2
    result = [None] * n_slots
    pending = [None] * n_slots
    nstep = [0]
    gather_bufs = None
    pusher = None
    peer = [None]
    gather_mode = a.gather if world > 1 else "none (1 GPU)"

    def step():
        nonlocal gather_bufs, pusher, gather_mode
        i = nstep[0] % n_slots
        nstep[0] += 1
        if pusher is not None:
            pusher.wait(i)
        if pending[i] is not None:     # the gather that still reads this result set
            pending[i].wait()
            pending[i] = None
        if world > 1 and gather_mode == "peer":
            # the result arrays of slot i live in rank 0's memory: this rank's kernels store them there directly
            if peer[0] is None:
                try:
                    from bfa_b200.aligner import result_arena_words
                    from bfa_b200.sharding import PeerArena
                    peer[0] = PeerArena(result_arena_words(B, bplan.max_stamps, True, True)["total"], dev, buffers=n_slots, dst=0)
                except Exception as e:   # noqa: BLE001
                    gather_mode = "root"
                    if rank == 0:
                        print(f"# peer-mapped result arrays unavailable ({e}); using copy-engine pushes", file=sys.stderr)
            if peer[0] is not None:
                r = dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=params, want_stamps=True, want_conf=True, plan=bplan, out=result[i],
                                    arena=peer[0].arena(i) if result[i] is None else None)
                result[i] = r
                return r
        r = dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=params, want_stamps=True, want_conf=True, plan=bplan, out=result[i])
        result[i] = r
        if world > 1:
            if gather_mode in ("root", "all") and pusher is None:   # first batch: the arena size is known now (collective call)
                try:
                    from bfa_b200.sharding import PushGather
                    pusher = PushGather(r.arena.numel(), dev, buffers=n_slots, dst=0 if gather_mode == "root" else None)
                except Exception as e:   # noqa: BLE001
                    gather_mode = "nccl"
                    if rank == 0:
                        print(f"# symmetric memory unavailable ({e}); using NCCL all_gather", file=sys.stderr)
            if pusher is not None:
                pusher.push(i, r.arena)
            else:
                if gather_bufs is None:   # stamps | conf | n_stamps | status | dp_final are one allocation: one collective
                    gather_bufs = [torch.empty(world * r.arena.numel(), dtype=r.arena.dtype, device=dev) for _ in range(n_slots)]
                pending[i] = dist.all_gather_into_tensor(gather_bufs[i], r.arena, async_op=True)
        return r

    def local_checksum(r):
        # what this rank computed, summed on this rank.  With peer-mapped arenas r.arena IS rank 0's memory, so the reference sum
        # comes from one more (untimed) alignment of the same batch into a zero-initialised local arena (the receive buffer
        # starts zeroed too, and the kernels leave unused stamp slots untouched)
        if peer[0] is None:
            return r.arena.to(torch.int64).sum().reshape(1)
        from bfa_b200.aligner import result_arena_words
        loc = dec.align_batch(lp, row_off, Ts, Cc, tgt32, Ns, params=params, want_stamps=True, want_conf=True, plan=bplan,
                              arena=torch.zeros(result_arena_words(B, bplan.max_stamps, True, True)["total"], dtype=torch.int32, device=dev))
        torch.cuda.synchronize()
        return loc.arena.to(torch.int64).sum().reshape(1)      # every step aligns the same batch: the same arrays, computed into LOCAL memory

    def drain():
        for i in range(n_slots):
            if pusher is not None:
                pusher.wait(i)
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Clock sampling starts here (nvidia-smi is already polling when the timed region begins).  Then the GPU is brought to its
    # sustained clocks: an idle B200 sits at 120 MHz and needs milliseconds of load to ramp up, far longer than W = 5 steps of
    # 0.15 ms -- so PREWARM_S seconds of untimed steps come first, then the contract's W warm-up steps, then (after the barrier)
    # the K timed steps, with no host-side pause in between.
    sampler = ClockSampler(local)
    sampler.start()
    # 0.15 s is enough for the SM clock to reach its maximum (measured: 1965 MHz with 0.05 s already); the roofline's denominator
    # is a burst figure (MEASURED_PEAKS.json: best of 10 copies), and a kernel timed alone is compared with that.  Under SUSTAINED
    # load these boxes run into their power cap: 1.5 s of pre-warm leaves the SM clock at 1837 MHz and the step at 0.148-0.149 ms
    # instead of 0.145 (profiles/r02_e_prewarm_sweep.txt); BFA_BENCH_PREWARM_S reproduces that.
    PREWARM_S = float(os.environ.get("BFA_BENCH_PREWARM_S", "0.15"))
    t_pre = time.perf_counter()
    n_pre = 0
    r = step(); drain(); torch.cuda.synchronize()
    while time.perf_counter() - t_pre < PREWARM_S:
        for _ in range(16):
            r = step()
        n_pre += 16
        drain()
        torch.cuda.synchronize()
    if world > 1 and (pusher is not None or peer[0] is not None):
        # untimed check of the gather: what landed in the receive buffer is what the ranks computed
        barrier()
        i_last = (nstep[0] - 1) % n_slots
        mine = local_checksum(result[i_last])
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, mine)
        if gather_mode == "all" or rank == 0:
            got = (pusher or peer[0]).recv[i_last].view(world, -1).to(torch.int64).sum(1)
            assert bool((got == sums).all()), "gather: receive buffer does not match the ranks' results"
        barrier()
    assert int((r.status[:B] & 7 != 0).sum()) == 0, "unexpected non-OK status on the synthetic workload"

    l0 = lib.bfa_launch_count()
    step(); drain(); torch.cuda.synchronize()
    one_kernel = (lib.bfa_launch_count() - l0) == 1     # the whole step is ONE kernel: the region's events are that kernel's events
    # More than one kernel per step: the dominant kernel is bracketed by CUDA events of its own (N=1: every step, N>1: every 4th
    # step - there the event records also serialise against the side-stream pushes).  One kernel per step: no events inside the
    # region, the kernel's average duration is the region's time / K.
    PROFILE_EVERY = 1 if world == 1 else 4
    lib.bfa_profile_enable(0 if one_kernel else (1 | (PROFILE_EVERY << 8)))
    lib.bfa_profile_read(None, None)
    nstep[0] = 0
    for _ in range(max(a.warmup, 3)):       # the W warm-up steps, immediately before the timed region
        step()
    drain()
    nstep[0] = 0                            # the timed steps use result sets / receive slots 0 .. K-1
    barrier()
    launches0 = lib.bfa_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    th0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    host_us_per_call = (time.perf_counter() - th0) / a.steps * 1e6     # how long the host needs to enqueue one step
    drain()                    # every push / gather has left inside the timed region ...
    e1.record()
    barrier()                  # ... and has landed (the barrier follows the last push of every rank)
    t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    dom_ms, dom_n = C.c_float(), C.c_int32()
    lib.bfa_profile_read(C.byref(dom_ms), C.byref(dom_n))
    lib.bfa_profile_enable(0)
    launches = lib.bfa_launch_count() - launches0
    clocks = sampler.stop(t_pre, t_end)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * T * a.steps / (ms / 1e3)
    gather_check = None
    if world > 1 and (pusher is not None or peer[0] is not None) and gather_mode in ("root", "peer"):
        # rank 0 now holds the timestamp arrays of every rank and every timed step: verify one (untimed)
        k = a.steps - 1
        mine = local_checksum(result[k % n_slots])
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, mine)
        if rank == 0:
            got = (pusher or peer[0]).recv[k % n_slots].view(world, -1).to(torch.int64).sum(1)
            gather_check = bool((got == sums).all())
            assert gather_check, "final gather: rank 0 does not hold the ranks' results of the last step"
        barrier()

    # ---- the same K steps once more (nothing differs when the step is one kernel; with the chain this run carries no per-kernel
    #      events).  Reported beside the contract's numbers, not instead of them.
    nstep[0] = 0
    barrier()
    plain_ms = time_steps(torch, step, a.steps, drain) * a.steps
    barrier()
    if world > 1:
        t = torch.tensor([plain_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        plain_ms = float(t.item())
    plain = {"ms_per_step": plain_ms / a.steps, "value": world * B * T * a.steps / (plain_ms / 1e3), "unit": "frames/s",
             "how": "the same K steps a second time, bfa_profile_enable(0)"}
    assert int((r.status[:B] & 7 != 0).sum()) == 0, "an utterance of the timed region was not finished (deferred / non-OK status)"

    # ---- roofline of the dominant kernel
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes(Ts, Ns, Cc)
    ctag = f"<{Cc},false>" if Cc in (66, 67, 17) else "<0,false> (run-time class count)"     # the compiled class counts (bfa_api.cu: direct_launch)
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
    # ---- the same kernel launched ALONE (events around every launch: nothing overlaps) and its fill phase by itself
    lib.bfa_profile_enable(1); lib.bfa_profile_read(None, None)
    nstep[0] = 0
    for _ in range(10):
        step()
    drain(); torch.cuda.synchronize()
    i_ms, i_n = C.c_float(), C.c_int32()
    lib.bfa_profile_read(C.byref(i_ms), C.byref(i_n)); lib.bfa_profile_enable(0)
    iso = i_ms.value / max(i_n.value, 1)
    roofline["isolated_launch_ms"] = iso
    roofline["isolated_launch_frac"] = (alg / (iso / 1e3) / 1e9 / peak) if iso > 0 else 0.0
    # north_star: ">= 70 % of the HBM roofline on the batched Viterbi fill": the same kernel with the measurement switch
    # BFA_FLAG_FILL_ONLY (rows streamed once, log-sum-exp, forward recursion, decision records written; no back-trace, no outputs)
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
                              "how": "same kernel, BFA_FLAG_FILL_ONLY (no back-trace, no outputs), 10 launches, CUDA events around each"}
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

    # ---- N=1 only: the other workloads the path is used on, and the reference-shaped call
    variants, api = None, None
    if world == 1 and not a.no_variants:
        variants = []

        def hinted(dec_, one):
            p = dec_._params(True, True, True)
            if one:
                p.reserved |= _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY | _cabi.FLAG_PIPELINED
            return p
        for Cv in (67, 17):        # the reference's real class counts (phoneme head 66 + blank, group head 16 + blank): run-time-C kernel
            lpv, tgv, _ = synth.planted_batch(B, T, N, Cv, seed=5000 + Cv, device=dev)
            decv = bfa_b200.AlignmentUtils(blank_id=Cv - 1, silence_id=0).viterbi_decoder
            w = dict(lp=lpv, row_off=torch.arange(B, dtype=torch.int64, device=dev) * T * Cv, Ts=Ts, Ns=Ns, tgt=tgv.to(torch.int32).reshape(-1).contiguous())
            variants.append(run_variant(torch, lib, decv, dev, f"metric shape at C={Cv} (blank {Cv - 1}), one kernel per step", w, Cv, hinted(decv, True), peak))
            del lpv, w
            torch.cuda.empty_cache()
        # both heads of one batch the way the shim enqueues them (core.py:900-922 aligns the phoneme head and the group head of every
        # batch): prepared first, then the two alignment kernels next to each other in the stream, the second one pipelined
        lp67, tg67, _ = synth.planted_batch(B, T, N, 67, seed=5067, device=dev)
        lp17, tg17, _ = synth.planted_batch(B, T, N, 17, seed=5017, device=dev)
        au67, au17 = bfa_b200.AlignmentUtils(blank_id=66, silence_id=0), bfa_b200.AlignmentUtils(blank_id=16, silence_id=0)
        lens_t, nl_t = torch.full((B,), T), torch.full((B,), N)
        h67 = au67.decode_alignments_prepare(lp67, tg67.cpu(), lens_t, nl_t)
        h17 = au17.decode_alignments_prepare(lp17, tg17.cpu(), lens_t, nl_t)

        def two_heads():
            au67.decode_alignments_enqueue(h67)
            au17.decode_alignments_enqueue(h17, after_sibling=True)
        for _ in range(4):
            two_heads()
        torch.cuda.synchronize()
        ms2 = time_steps(torch, two_heads, 10)
        ok2 = int((h67["r"].status[:B] & 7 != 0).sum()) == 0 and int((h17["r"].status[:B] & 7 != 0).sum()) == 0
        alg2 = algorithmic_bytes(Ts, Ns, 67) + algorithmic_bytes(Ts, Ns, 17)
        variants.append({"name": "both heads of the metric shape per step (C=67 then C=17, prepared first, second kernel pipelined): what core.py:900-922 costs",
                         "B": B, "C": [67, 17], "frames": B * T, "ms_per_step": ms2, "value": B * T / (ms2 / 1e3), "unit": "frames/s (utterance frames, both heads aligned)",
                         "launches_per_step": 2.0, "all_finished": ok2,
                         "roofline": {"algorithmic_bytes": alg2, "achieved": alg2 / (ms2 / 1e3) / 1e9, "frac": alg2 / (ms2 / 1e3) / 1e9 / peak, "unit": "GB/s",
                                      "of": "both kernels, CUDA events around 10 steps"},
                         "items": {"exact_kernel": 0, "window_24_40_64": [0, 0, 0]}})
        # the same C = 67 head from un-normalised LOGITS (core.py:898-899 not run): one kernel (bfa_align_batch_logits), next to what the
        # reference's order of operations costs here -- a log-softmax pass over [B, T, C] (read + write) and then the kernel
        from bfa_b200.aligner import log_softmax_rows
        lg67 = lp67 + 3.0                                                  # any per-row shift: the kernel never sees normalised rows
        tg67_32 = tg67.to(torch.int32).reshape(-1).contiguous()
        p67 = hinted(au67.viterbi_decoder, True)
        plan67 = au67.viterbi_decoder.plan_batch(Ts, Ns, 67, params=p67, device=dev)
        rl = [None, None]

        def from_logits():
            rl[0] = au67.viterbi_decoder.align_batch(lg67, row_off67, Ts, 67, tg67_32, Ns, params=p67, plan=plan67, out=rl[0], logits=True)

        p67s = au67.viterbi_decoder._params(True, True, True)              # not BFA_FLAG_PIPELINED: its rows are written by the kernel right before it
        p67s.reserved |= _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY
        plan67s = au67.viterbi_decoder.plan_batch(Ts, Ns, 67, params=p67s, device=dev)

        def softmax_then_align():
            rl[1] = au67.viterbi_decoder.align_batch(log_softmax_rows(lg67), row_off67, Ts, 67, tg67_32, Ns, params=p67s, plan=plan67s, out=rl[1])
        row_off67 = torch.arange(B, dtype=torch.int64, device=dev) * (T * 67)
        for f in (from_logits, softmax_then_align):
            for _ in range(4):
                f()
        torch.cuda.synchronize()
        ms_l, ms_s = time_steps(torch, from_logits, 10), time_steps(torch, softmax_then_align, 10)
        same = bool(torch.equal(rl[0].frame_ph, rl[1].frame_ph)) and int((rl[0].status[:B] & 7 != 0).sum()) == 0
        alg67 = algorithmic_bytes(Ts, Ns, 67)
        variants.append({"name": "metric shape at C=67 from un-normalised logits, one kernel per step (bfa_align_batch_logits: F.log_softmax of core.py:898-899 never run)",
                         "B": B, "C": 67, "frames": B * T, "ms_per_step": ms_l, "value": B * T / (ms_l / 1e3), "unit": "frames/s",
                         "launches_per_step": 1.0, "all_finished": same,
                         "log_softmax_pass_then_kernel_ms": ms_s, "frames_identical_to_that": same,
                         "roofline": {"algorithmic_bytes": alg67 + 4 * B * T, "achieved": (alg67 + 4 * B * T) / (ms_l / 1e3) / 1e9,
                                      "frac": (alg67 + 4 * B * T) / (ms_l / 1e3) / 1e9 / peak, "unit": "GB/s",
                                      "of": "the kernel (rows read once, frame labels + row_lse written), CUDA events around 10 steps"},
                         "items": {"exact_kernel": 0, "window_24_40_64": [0, 0, 0]}})
        del lg67, rl
        del lp67, lp17, h67, h17
        torch.cuda.empty_cache()
        # SIL at every 10th target (what punctuation does to real targets): silence anchoring really runs -- row statistics for the
        # silence scan (a second read of the rows), planner, segments through the list-mode banded kernel, stamp kernel
        lpv, tgv, _ = synth.planted_batch(B, T, N, Cc, seed=6001, peak=10.0, sil_every=10, sil_frames=15, device=dev)
        w = dict(lp=lpv, row_off=row_off, Ts=Ts, Ns=Ns, tgt=tgv.to(torch.int32).reshape(-1).contiguous())
        variants.append(run_variant(torch, lib, dec, dev, "metric shape with silence_id at every 10th target (anchoring on, no hints, full chain)", w, Cc,
                                    hinted(dec, False), peak))
        # the same batch as un-normalised logits: the chain runs on them as they are (the silence pass also leaves row_lse), next to
        # what the reference's order costs here: a log-softmax pass over [B, T, C], then the chain
        from bfa_b200.aligner import log_softmax_rows as _lsr
        lgv = lpv + 2.0
        pl_ = hinted(dec, False)
        plan_l = dec.plan_batch(Ts, Ns, Cc, params=pl_, device=dev)
        rr = [None, None]

        def sil_from_logits():
            rr[0] = dec.align_batch(lgv, row_off, Ts, Cc, w["tgt"], Ns, params=pl_, plan=plan_l, out=rr[0], logits=True)

        def sil_softmax_first():
            rr[1] = dec.align_batch(_lsr(lgv), row_off, Ts, Cc, w["tgt"], Ns, params=pl_, plan=plan_l, out=rr[1])
        for f in (sil_from_logits, sil_softmax_first):
            for _ in range(3):
                f()
        torch.cuda.synchronize()
        ms_a, ms_b = time_steps(torch, sil_from_logits, 10), time_steps(torch, sil_softmax_first, 10)
        same_f = float((rr[0].frame_ph == rr[1].frame_ph).float().mean())
        algs = algorithmic_bytes(Ts, Ns, Cc) + 4 * B * T
        variants.append({"name": "metric shape with silence_id at every 10th target, from un-normalised logits (full chain on the logits, silence pass leaves row_lse)",
                         "B": B, "C": Cc, "frames": B * T, "ms_per_step": ms_a, "value": B * T / (ms_a / 1e3), "unit": "frames/s",
                         "launches_per_step": 7.0, "log_softmax_pass_then_chain_ms": ms_b, "frame_labels_equal_fraction": same_f,
                         "roofline": {"algorithmic_bytes": algs, "achieved": algs / (ms_a / 1e3) / 1e9, "frac": algs / (ms_a / 1e3) / 1e9 / peak,
                                      "unit": "GB/s", "of": "the whole step (every kernel of the call), CUDA events around 10 steps"},
                         "items": {"exact_kernel": 0, "window_24_40_64": [0, 0, 0]}})
        del lgv, rr
        psil = hinted(dec, False)
        psil.reserved |= _cabi.FLAG_NO_DIRECT        # what the facade sets when the host-side targets (nearly) all hold silence_id
        variants.append(run_variant(torch, lib, dec, dev, "metric shape with silence_id at every 10th target, one-kernel pass skipped (BFA_FLAG_NO_DIRECT)", w, Cc,
                                    psil, peak))
        del lpv, w
        torch.cuda.empty_cache()
        for n in (2, 3, 4):        # BASELINE.json configs[1..3]
            w = synth.baseline_config(n, C=Cc, device=dev)
            variants.append(run_variant(torch, lib, dec, dev, "BASELINE config " + w["name"], w, Cc, hinted(dec, n == 2), peak))
            del w
            torch.cuda.empty_cache()
        # ---- the call core.py makes: AlignmentUtils.decode_alignments on a [B, T, C] CUDA tensor with host-side targets, Python
        #      lists of tuples out (forced_alignment.py:856-910)
        lens, nl = torch.full((B,), T), torch.full((B,), N)
        tgt_host = tgt.cpu()
        for _ in range(2):
            out = au.decode_alignments(lp, true_seqs=tgt_host, pred_lens=lens, true_seqs_lens=nl, with_confidence=True)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            out = au.decode_alignments(lp, true_seqs=tgt_host, pred_lens=lens, true_seqs_lens=nl, with_confidence=True)
        dt = (time.perf_counter() - t0) / reps
        api = {"call": "AlignmentUtils.decode_alignments(log_probs[B,T,C] on the GPU, host targets, with_confidence=True) -> list[B] of lists of tuples",
               "ms_per_call": dt * 1e3, "value": B * T / dt, "unit": "frames/s", "stamps_returned": int(sum(len(x) for x in out))}

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
        # the fabric's ceiling for this step: nothing but the same bytes copied (pinned H2D + D2H on two streams), all ranks at once
        d_in = torch.empty_like(lp)
        d_out = torch.empty(B * T * 2 + B * ms_stamps * 5 + 3 * B, dtype=torch.int32, device=dev)
        h_out = torch.empty(d_out.shape, dtype=torch.int32, pin_memory=True)
        s2 = torch.cuda.Stream()
        barrier()
        t1 = time.perf_counter()
        for _ in range(ke):
            d_in.copy_(lp_h, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt_copy = time.perf_counter() - t1
        if world > 1:
            t = torch.tensor([dt, dt_copy], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_copy = float(t[0].item()), float(t[1].item())
        h2d = lp_h.numel() * 4 + tgt_h.numel() * 4 + B * 8 + B * 4 + 2 * (B + 1) * 8
        d2h = sum(v.nbytes for v in out.values() if hasattr(v, 'nbytes') and v is not out.get('frame_off'))
        e2e = {"value": world * B * T * ke / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": ke, "ms_per_step": dt / ke * 1e3, "api": "bfa_align_batch_host (pinned host buffers, 512-utterance chunks, 2 streams)",
               "copy_only_ms_per_step": dt_copy / ke * 1e3,
               "copy_only_note": "the same H2D + D2H bytes with plain pinned cudaMemcpyAsync on two streams, every rank at once: what the host "
                                 "fabric (PCIe / host memory) allows; e2e runs at this rate when its ms_per_step is close to it"}
        lib.bfa_host_release()
        del d_in, d_out

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n = a.cpu_sample or B
        cpu_baseline, _ = cpu_port_arm(a, n, 8, 1)    # ~1 s wall on 16 threads = ~17 core-seconds of CPU work
        try:
            ref, _ = python_reference_arm(a, 1, 2, 0)   # the unmodified Python reference, one utterance per core and pass
        except Exception as e:   # noqa: BLE001
            ref = {"error": str(e)[:200]}
        cpu_baseline["python_reference"] = ref

    if rank == 0:
        gather_txt = {"none (1 GPU)": "none (1 GPU)",
                      "peer": "the alignment kernels of every rank store their packed result arrays of every step straight into rank 0's memory "
                              "(peer-mapped stores over NVLink inside the kernel, torch symmetric memory, sharding.PeerArena(dst=0)): no copy, no side "
                              "stream, nothing between two launches; complete when the last kernel is, verified after the timed region",
                      "root": "every rank streams its packed result arrays of every step to rank 0 (copy-engine P2P writes over NVLink on a side stream, "
                              "sharding.PushGather(dst=0)); completed inside and verified after the timed region",
                      "all": "every rank pushes its packed result arrays to ALL peers each step (stress variant)",
                      "nccl": "NCCL all_gather_into_tensor per step"}[gather_mode]
        launch_txt = ("one kernel per step (BFA_FLAG_DIRECT_ONLY" + (" | BFA_FLAG_PIPELINED" if params.reserved & _cabi.FLAG_PIPELINED else "") +
                      "): the targets hold no silence_id and every utterance is a plain stride-4 problem (host-side knowledge); statuses checked "
                      "after the timed region") if one_kernel else "full chain"
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
               "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "impl": "b200",
               "config": workload_config(a, {"gather": gather_txt, "sharding": f"{world} rank(s) x {B} utterances per step, no data-path collective",
                                             "launch": launch_txt,
                                             "prewarm": f"{n_pre} untimed steps ({PREWARM_S} s) bring the GPU to its sustained clocks before the W warm-up steps"}),
               "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "uninstrumented_step": plain, "host_enqueue_us_per_step": host_us_per_call, "prewarm_steps": n_pre,
               "variants": variants, "reference_api_call": api, "final_gather_verified": gather_check}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
def corpus_main(a, world, rank, local, dev, lib):
    """BASELINE config 5: a corpus of `--corpus` utterances of the config-2 shape, sharded over the ranks by sharding.shard_ranges,
    aligned chunk by chunk on each rank (posteriors generated on the device, never moved between GPUs), the timestamp arrays
    streamed to rank 0 as they are produced -- ONE gather to one place, complete when the last chunk's push has landed."""
    import torch
    import torch.distributed as dist
    import bfa_b200
    from bfa_b200 import _cabi, synth
    from bfa_b200.sharding import shard_ranges

    n_total, T, N, Cc, CH = a.corpus, a.T, a.N, a.C, a.batch
    # equal utterances: the work-balanced ranges are equal counts, rounded to whole chunks (the last rank takes the remainder)
    per = (n_total // world + CH - 1) // CH * CH
    ranges = [(min(r * per, n_total), min((r + 1) * per, n_total)) for r in range(world)]
    ranges[-1] = (ranges[-1][0], n_total)
    assert shard_ranges([T] * 8, [N] * 8, 2) == [(0, 4), (4, 8)]
    s0, e0 = ranges[rank]
    n_chunks = (e0 - s0 + CH - 1) // CH
    chunks_per_rank = [(e - s + CH - 1) // CH for s, e in ranges]
    n_chunks_max = max(chunks_per_rank)
    au = bfa_b200.AlignmentUtils(blank_id=Cc - 1, silence_id=0)
    dec = au.viterbi_decoder
    params = dec._params(True, True, True)
    params.reserved |= _cabi.HINT_NO_SIL | _cabi.FLAG_DIRECT_ONLY | _cabi.FLAG_PIPELINED
    # a pool of distinct chunks generated on the device (the corpus is synthetic: chunk c of a rank is pool entry c % POOL);
    # generating 158 GB of posteriors is not what is being measured, reading 158 GB of posteriors from HBM is
    POOL = 4
    pool = []
    for k in range(POOL):
        lpk, tgk, _ = synth.planted_batch(CH, T, N, Cc, seed=9000 + 100 * rank + k, device=dev)
        pool.append((lpk, tgk.to(torch.int32).reshape(-1).contiguous()))
    row_off = torch.arange(CH, dtype=torch.int64, device=dev) * (T * Cc)
    Ts, Ns = [T] * CH, [N] * CH
    plan = dec.plan_batch(Ts, Ns, Cc, params=params, device=dev)
    results = [None] * max(n_chunks_max, 1)
    r0 = dec.align_batch(pool[0][0], row_off, Ts, Cc, pool[0][1], Ns, params=params, plan=plan)
    words = r0.arena.numel()
    pusher = None
    peer = None
    if world > 1:
        from bfa_b200.sharding import PeerArena, PushGather
        if a.gather == "peer":
            try:
                peer = PeerArena(words, dev, buffers=n_chunks_max, dst=0)
            except Exception as e:   # noqa: BLE001
                if rank == 0:
                    print(f"# peer-mapped result arrays unavailable ({e}); using copy-engine pushes", file=sys.stderr)
        if peer is None:
            pusher = PushGather(words, dev, buffers=n_chunks_max, dst=0)
    for c in range(n_chunks):               # result sets allocated up front: local arenas (pushed) or this rank's slots in rank 0's memory
        results[c] = dec.align_batch(pool[c % POOL][0], row_off, Ts, Cc, pool[c % POOL][1], Ns, params=params, plan=plan,
                                     arena=peer.arena(c) if peer is not None else None)
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.4:   # clocks
        for k in range(8):
            r0 = dec.align_batch(pool[k % POOL][0], row_off, Ts, Cc, pool[k % POOL][1], Ns, params=params, plan=plan, out=r0)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t0 = time.perf_counter()
    ev[0].record()
    for c in range(n_chunks):
        lpk, tgk = pool[c % POOL]
        results[c] = dec.align_batch(lpk, row_off, Ts, Cc, tgk, Ns, params=params, plan=plan, out=results[c])
        if pusher is not None:
            pusher.push(c, results[c].arena)
    ev[1].record()                          # the last alignment kernel of this rank
    if pusher is not None:
        for c in range(n_chunks):
            pusher.wait(c)
    ev[2].record()                          # ... and its last push has left
    if world > 1:
        dist.barrier()                      # every rank's pushes have landed on rank 0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    align_ms, tail_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    ok = all(int((results[c].status[:CH] & 7 != 0).sum()) == 0 for c in range(0, n_chunks, max(1, n_chunks // 4)))
    tt = torch.tensor([align_ms, tail_ms, wall * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    verified = None
    if world > 1:                           # rank 0 holds every rank's arrays: check the last chunk every rank has
        last = min(chunks_per_rank) - 1
        if peer is not None:                # the arena is rank 0's memory: the reference sum comes from the same chunk aligned into local memory
            loc = dec.align_batch(pool[last % POOL][0], row_off, Ts, Cc, pool[last % POOL][1], Ns, params=params, plan=plan,
                                  arena=torch.zeros(words, dtype=torch.int32, device=dev))
            torch.cuda.synchronize()
            mine_sum = loc.arena.to(torch.int64).sum().reshape(1)
        else:
            mine_sum = results[last].arena.to(torch.int64).sum().reshape(1)
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, mine_sum)
        if rank == 0:
            got = (peer or pusher).recv[last].view(world, -1).to(torch.int64).sum(1)
            verified = bool((got == sums).all())
    if rank == 0:
        frames = n_total * T
        out = {"metric": "BASELINE config 5: 1M-utterance synthetic corpus sharded across the GPUs of one box, gather of timestamp arrays",
               "corpus_utterances": n_total, "n_gpus": world, "chunk_utterances": CH, "chunks_per_rank": chunks_per_rank,
               "align_ms_max_over_ranks": float(tt[0]), "gather_tail_ms_after_last_kernel": float(tt[1]), "wall_ms_incl_final_barrier": float(tt[2]),
               "value": frames / (float(tt[2]) / 1e3), "unit": "frames/s", "value_device_timed": frames / ((float(tt[0]) + float(tt[1])) / 1e3),
               "gathered_bytes_on_rank0": int(words * 4 * sum(chunks_per_rank)) if world > 1 else 0,
               "statuses_ok": ok, "gather_verified": verified,
               "how": "every rank aligns its contiguous share chunk by chunk (one kernel per chunk, pipelined launches)"
                      + (" and its kernels store each chunk's packed result arrays straight into rank 0's memory (peer-mapped stores over NVLink)"
                         if peer is not None else " and streams each chunk's packed result arrays to rank 0 with copy-engine P2P writes")
                      + "; the only synchronisation is the barrier at the end",
               "data": f"synthetic, {POOL} distinct chunks per rank generated on the device and cycled (the whole corpus would be {frames * Cc * 4 / 1e9:.0f} GB of posteriors); "
                       "utterances beyond the corpus size in the last chunk of a rank are aligned too (whole chunks)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
