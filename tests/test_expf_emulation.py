"""The CUDA library evaluates exp() in the exp-weighted stages (WM, censusGrad's gradient term, gen_vm_from2vm_exp on
materialised volumes) with a restatement of the host libm's expf (smd_expf_host, csrc/common.cuh) so that those stages
match the CPU reference bit for bit.  oracle/expf_check.c holds the same operations for the host; this test checks them
against THIS machine's libm on a dense sample (the exhaustive run over [-320, 100], 2.25e9 floats, is
`gcc -O2 -ffp-contract=off -mfma oracle/expf_check.c -lm && ./a.out`: 0 mismatches with glibc 2.39).  The GPU side of
the claim is tests/test_gpu_parity.py::test_combine_exp_bit_exact."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    so = os.path.join(ROOT, "oracle", "libexpfcheck.so")
    src = os.path.join(ROOT, "oracle", "expf_check.c")
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libexpfcheck.so"])
    L = C.CDLL(so)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.sm_expf_both.argtypes = [f32p, C.c_long, f32p, f32p]
    L.sm_expf_both.restype = None
    return L


def test_restated_expf_equals_libm_on_a_dense_sample():
    L = _lib()
    rng = np.random.default_rng(0)
    # every 97th float bit pattern of [-320, 0) and (0, 100], plus the WM / censusGrad argument shapes
    neg = np.arange(0x80000000, np.float32(-320.0).view(np.uint32) + 1, 97, dtype=np.uint64).astype(np.uint32).view(np.float32)
    pos = np.arange(0, np.float32(100.0).view(np.uint32) + 1, 97, dtype=np.uint64).astype(np.uint32).view(np.float32)
    c2 = rng.integers(0, 3 * 255 * 255 + 1, 2_000_000).astype(np.float32)
    sp = rng.integers(0, 163, 2_000_000).astype(np.float32)
    wm = (-c2 / np.float32(625.0)) - (sp / np.float32(81.0))
    special = np.array([0.0, -0.0, -103.97, -103.98, -104.0, -87.3, -88.0, 88.72, 88.73, 89.0, np.inf, -np.inf, np.nan],
                       np.float32)
    x = np.ascontiguousarray(np.concatenate([neg, pos, wm.astype(np.float32), special]))
    a, b = np.empty_like(x), np.empty_like(x)
    L.sm_expf_both(x, x.size, a, b)
    bad = a.view(np.uint32) != b.view(np.uint32)
    bad &= ~(np.isnan(a) & np.isnan(b))
    assert not bad.any(), (x[bad][:5], a[bad][:5], b[bad][:5])
