"""CPU suite: the C-ABI library loads and exports every symbol include/*.h declares; the ctypes table in
mystereomatching_b200/capi.py covers exactly the same set; no compute call is made (no GPU here)."""
import ctypes
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(sm_[a-z0-9_]+|SM_[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n.startswith("sm_")))


def test_header_declares_something():
    d = _declared()
    assert len(d) > 40 and "sm_pipeline_run" in d and "sm_tree_filter" in d


def test_every_declared_symbol_is_exported():
    from mystereomatching_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "build the library first (__graft_entry__.build())"
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    missing = [n for n in _declared() if n not in exported]
    assert not missing, missing
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in _declared():
        getattr(lib, n)


def test_ctypes_table_matches_header():
    from mystereomatching_b200 import capi
    assert sorted(capi.SIGNATURES) == _declared()


def test_params_default_and_struct_layout():
    from mystereomatching_b200 import capi
    p = capi.default_params(63)
    assert p.numDisparities == 64 and p.censusFunc == 3 and p.sgm_paths == 4
    assert (p.cbca_crossL, p.cbca_crossL_out, p.cbca_cTresh, p.cbca_cTresh_out) == (17, 34, 20, 6)
    assert p.adTrunc == 1000.0 and p.lamAD == 10.0 and p.lamCen == 30.0
    assert p.DISP_OCC == -32 and p.DISP_MIS == -48 and p.crossScaleLambda < 0
    assert p.costcalculation == 0 and p.cg_lamCen == 13.0 and p.cg_lamG == 1.0 and p.gradTrunc == 500.0
    assert p.pyramidLevels == 1
    # sizeof check: the header struct is 41 4-byte fields
    assert ctypes.sizeof(capi.SmParams) == 41 * 4


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mystereomatching_b200 import capi
    assert capi.lib().sm_device_count() == 0
    h = ctypes.c_void_p()
    assert capi.lib().sm_ctx_create(ctypes.byref(h), 0, None) != 0
    assert b"no CUDA device" in capi.lib().sm_last_error()
    with pytest.raises(capi.SmError):
        capi.Ctx(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mystereomatching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "orc_" not in txt, f
