"""Recipe: stage the UNMODIFIED reference implementation of the path under oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

The reference's path is two plain Python files over torch (bournemouth_aligner/forced_alignment.py, utils.py).  The package
itself cannot be imported (its __init__ pulls in phonemizer / espeak), and /root/reference does not exist on the GPU box, so
this recipe copies the two files byte for byte from where they lie under /root/reference into oracle/_ref/ (git-ignored: they
never enter the history; NOT gpurun-ignored: they travel to the GPU box like a built .so).  `load()` then imports them by path,
exactly like tests/golden/make_golden.py does in the build container.

Only bench.py's reference arm (`--impl reference`, `cpu_baseline.kind == "reference"`) and oracle/validate_against_reference.py
use this; nothing under bournemouth-forced-aligner_b200/ does, and nothing is ever copied into tracked files.
"""
from __future__ import annotations

import hashlib
import importlib.util
import shutil
import sys
from pathlib import Path

_HERE = Path(__file__).resolve().parent
REF_DIR = _HERE / "_ref"
SOURCE = Path("/root/reference/bournemouth_aligner")
FILES = ("forced_alignment.py", "utils.py")


def stage(force: bool = False) -> bool:
    """Copy the reference files into oracle/_ref/ when /root/reference is present.  Returns True when the staged copy exists."""
    if SOURCE.is_dir():
        REF_DIR.mkdir(parents=True, exist_ok=True)
        for f in FILES:
            src, dst = SOURCE / f, REF_DIR / f
            if force or not dst.exists() or dst.read_bytes() != src.read_bytes():
                shutil.copyfile(src, dst)
        (REF_DIR / "SHA256").write_text("".join(f"{hashlib.sha256((REF_DIR / f).read_bytes()).hexdigest()}  {f}\n" for f in FILES))
    return available()


def available() -> bool:
    return all((REF_DIR / f).exists() for f in FILES)


def load():
    """(forced_alignment module, utils module) of the staged, unmodified reference."""
    mods = []
    for f in FILES:
        name = "bfa_reference_" + f[:-3]
        if name in sys.modules:
            mods.append(sys.modules[name]); continue
        spec = importlib.util.spec_from_file_location(name, REF_DIR / f)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


def _noop(_):
    """Pool warm-up: the worker imports torch and the staged reference once."""
    import torch
    torch.set_num_threads(1)
    load()
    return 0


def _worker(args):
    """One process = one host core: the reference's own batch entry on a slice of the sample."""
    import time
    import torch
    lp, tgt, T, N, blank, with_conf = args
    torch.set_num_threads(1)
    fa, ut = load()
    au = fa.AlignmentUtils(blank_id=blank, silence_id=0, silence_anchors=10, ignore_noise=True, truly_forced=True)   # core.py:256-257
    B = lp.shape[0]
    t0 = time.perf_counter()
    frames = au.decode_alignments(lp, true_seqs=tgt, pred_lens=torch.full((B,), T), true_seqs_lens=torch.full((B,), N),
                                  forced_alignment=True, boost_targets=True, enforce_minimum=True)
    if with_conf:
        for b in range(B):
            ut._calculate_confidences(lp[b], [f + (False,) for f in frames[b]])          # core.py:936-937
    return time.perf_counter() - t0, frames


def time_reference(lp, tgt, T, N, blank, n_procs, with_conf=True):
    """Run the staged reference on lp [B,T,C] / tgt [B,N] split over n_procs worker processes (torch threads = 1 each).
    Returns (wall seconds of the slowest worker, list of per-utterance stamp lists)."""
    import torch.multiprocessing as mp
    B = lp.shape[0]
    n_procs = max(1, min(n_procs, B))
    bounds = [B * i // n_procs for i in range(n_procs + 1)]
    jobs = [(lp[bounds[i]:bounds[i + 1]].clone(), tgt[bounds[i]:bounds[i + 1]].clone(), T, N, blank, with_conf) for i in range(n_procs)]
    if n_procs == 1:
        res = [_worker(jobs[0])]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Pool(n_procs) as pool:
            res = pool.map(_worker, jobs)
    return max(r[0] for r in res), [f for r in res for f in r[1]]


if __name__ == "__main__":
    print("staged" if stage(force="--force" in sys.argv) else "reference not available (no /root/reference, nothing staged)")
