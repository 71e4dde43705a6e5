"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/ndt2d.h
declares, struct layouts agree with the header, and the product fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ndt2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ndt2d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gtsam_ndt_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ndt2d.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.ndt2d_version() == 200


def test_struct_layouts_match_header_and_oracle():
    from gtsam_ndt_b200 import _lib, matcher
    import oracle
    assert C.sizeof(_lib.Params) == 11 * 8 + 4 * 4 == C.sizeof(oracle.Params)
    assert [f[0] for f in _lib.Params._fields_] == [f[0] for f in oracle.Params._fields_]
    assert matcher.RESULT_DTYPE == oracle.RESULT_DTYPE
    p, q = _lib.Params(), oracle.Params()
    _lib.load().ndt2d_default_params(C.byref(p))
    oracle.lib().oracle_default_params(C.byref(q))
    assert bytes(p) == bytes(q)            # same defaults (SPEC.md section 1)
    assert (p.eig_ratio, p.min_points, p.max_iterations, p.lambda_down) == (0.01, 3, 30, 5.0)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gtsam_ndt_b200 as g
    with pytest.raises(g.NdtError, match="no CPU fallback"):
        g.NdtMatcher2D([0.5])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gtsam_ndt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ndt2d_oracle" not in txt, f
    hpp = open(os.path.join(ROOT, "include", "ndt2d.h")).read()
    assert "oracle" not in hpp.lower().replace("spec oracle", "")
