"""Build libbfa_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "lib" / "libbfa_b200.so"
SOURCES = ["bfa_api.cu", "bfa_host.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--threads", "2"]


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "bfa_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: Path = None) -> Path:
    if out is None and not force and not _stale():
        return LIB
    lib = out or LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(lib), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:      # A/B build: python build.py --variant NAME DEFINE [DEFINE...] -> lib/libbfa_b200_NAME.so
        i = sys.argv.index("--variant")
        print(build(force=True, defines=tuple(sys.argv[i + 2:]), out=HERE / "lib" / f"libbfa_b200_{sys.argv[i + 1]}.so"))
    elif "--phase-prof" in sys.argv:   # development build with per-phase cycle counters in the banded kernel
        print(build(force=True, defines=("BFA_PHASE_PROF",), out=HERE / "lib" / "libbfa_b200_prof.so"))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
