#!/usr/bin/env python
"""tests/golden/post.npz: utils.convert_to_ms of the UNMODIFIED reference on seeded stamp lists (build container only)."""
import importlib.util
from pathlib import Path
import numpy as np

REF = Path("/root/reference/bournemouth_aligner/utils.py")
spec = importlib.util.spec_from_file_location("bfa_ref_utils", REF)
u = importlib.util.module_from_spec(spec); spec.loader.exec_module(u)
rng = np.random.default_rng(5)
out = {}
for case in range(12):
    n = int(rng.integers(0, 40))
    T = int(rng.integers(1, 3000)) if case != 3 else 0
    starts = np.sort(rng.integers(0, max(T, 1), n))
    ends = starts + rng.integers(1, 30, n)
    tl = int(rng.integers(3, 7)) if case % 4 else 6
    stamps = [tuple([int(rng.integers(1, 66)), int(s), int(e), int(i), bool(rng.integers(0, 2)), float(rng.random())][:tl])
              for i, (s, e) in enumerate(zip(starts, ends))]
    off = float(rng.random() * 100) if case % 3 else 0.0
    wav_len = int(rng.integers(1000, 16000 * 60)); sr = [16000, 22050, 44100][case % 3]
    got = u.convert_to_ms(stamps, T, off, wav_len, sr)
    out[f"c{case}/args"] = np.array([T, off, wav_len, sr, tl], np.float64)
    out[f"c{case}/stamps"] = np.array([[float(x) for x in s] + [0.0] * (6 - len(s)) for s in stamps], np.float64).reshape(n, 6)
    out[f"c{case}/ms"] = np.array([[g[6], g[7]] for g in got], np.float64).reshape(n, 2)
    out[f"c{case}/full"] = np.array([[float(x) for x in g] for g in got], np.float64).reshape(n, 8)
np.savez_compressed(Path(__file__).resolve().parent / "post.npz", **out)
print("wrote post.npz", len(out))
