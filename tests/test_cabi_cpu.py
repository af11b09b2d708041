"""CPU: the C-ABI library loads and exports every symbol include/bfa_b200.h declares; struct layouts and
constants agree between the header, the library and the ctypes binding.  No compute calls."""
import ctypes as C
import re
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def _header_functions():
    text = (ROOT / "include" / "bfa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bfa_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from bfa_b200 import _cabi
    lib = _cabi.lib()
    names = _header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bfa_b200.h but not exported"
    assert set(names) == set(_cabi.exported_symbols())


def test_version_and_strerror():
    from bfa_b200 import _cabi
    lib = _cabi.lib()
    assert lib.bfa_version() == 100
    assert lib.bfa_strerror(0) == b"ok"
    assert b"workspace" in lib.bfa_strerror(_cabi.BFA_E_WORKSPACE)


def test_params_layout_and_defaults():
    from bfa_b200 import _cabi
    lib = _cabi.lib()
    assert lib.bfa_sizeof_params() == C.sizeof(_cabi.BfaParams) == 64
    p = _cabi.default_params(66, 0)
    assert (p.blank_id, p.silence_id, p.silence_anchors, p.ignore_noise, p.truly_forced) == (66, 0, 10, 1, 1)
    assert (p.boost_targets, p.enforce_minimum, p.max_blanks, p.boundary_pad, p.min_speech_frames) == (1, 1, 10, 3, 20)
    assert p.boost_factor == 5.0 and p.neg_inf == -1000.0 and p.sub_boost == 5.0
    assert np.float32(p.min_log_prob) == np.log(np.float32(1e-8))
    assert _cabi.default_params(16, None).silence_id == -1


def test_oracle_and_product_params_agree():
    """The oracle's parameter block is the same POD (so parity tests vary the same constants)."""
    from bfa_b200 import _cabi
    from oracle import oracle as orc
    a, b = _cabi.default_params(66, 0), orc.params(66, 0)
    assert bytes(a) == bytes(b)


def test_invalid_arguments_are_rejected_without_a_gpu():
    from bfa_b200 import _cabi
    lib = _cabi.lib()
    p = _cabi.default_params(66, 0)
    s = _cabi.BfaShape(1, 67, 10, 2, 10, 12, 0)
    rc = lib.bfa_align_batch(C.byref(p), C.byref(s), *([None] * 13), None, 0, None)
    assert rc == _cabi.BFA_E_INVALID
    assert lib.bfa_confidence_batch(1, 67, None, None, None, None, None, 4, None, None) == _cabi.BFA_E_INVALID


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (CPU fallback would void parity claims)."""
    pkg = ROOT / "bournemouth-forced-aligner_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + [ROOT / "bfa_b200" / "__init__.py"]:
        txt = f.read_text()
        assert "oracle" not in txt.replace("no dependency on oracle/", ""), f
