"""The file side of the reference's driver: mystereomatching_b200/host/sm_io.{h,cpp} against what
cv2 4.13's imread / imwrite do with the same files (tests/golden/make_io_golden.py wrote the files and io_ref.npz;
cv2 is not imported here).  Reference: main_.cpp:31-39 (dataset table), :85-129 (the reads and DT.convertTo),
stereoMatching.h:2005-2110 (saveDispMap).  Host code only: runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mystereomatching_b200", "libsm_io.so")
GOLD = os.path.join(ROOT, "tests", "golden")
IO = os.path.join(GOLD, "io")


@pytest.fixture(scope="module")
def lib():
    assert os.path.exists(LIB), "run __graft_entry__.build()"
    return C.CDLL(LIB)


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "io_ref.npz"))


def _imread(lib, path, flag):
    h, w, c = C.c_int(), C.c_int(), C.c_int()
    if lib.smio_image_dims(path.encode(), C.byref(h), C.byref(w), C.byref(c)) != 0:
        return None
    out = np.empty((h.value, w.value, 3) if flag else (h.value, w.value), np.uint8)
    assert lib.smio_imread(path.encode(), flag, out.ctypes.data_as(C.c_void_p), h.value, w.value) == 0
    return out


@pytest.mark.parametrize("name", ["rgb8.png", "gray8.png", "gray16.png", "rgba8.png", "rgb16.png", "rgb8.ppm", "gray8.pgm",
                                  "smooth.png"])
def test_imread_equals_cv2(lib, ref, name):
    for flag in (1, 0):
        got = _imread(lib, os.path.join(IO, name), flag)
        assert got is not None
        assert np.array_equal(got, ref[f"{name}:{flag}"]), (name, flag)


def test_missing_file_gives_empty_image(lib):
    assert _imread(lib, os.path.join(IO, "nope.png"), 1) is None


def test_pfm_roundtrip_and_cv2_orientation(lib, ref, tmp_path):
    h, w = C.c_int(), C.c_int()
    p = os.path.join(IO, "gt.pfm")
    assert lib.smio_pfm_dims(p.encode(), C.byref(h), C.byref(w)) == 0
    out = np.empty((h.value, w.value), np.float32)
    assert lib.smio_read_pfm(p.encode(), out.ctypes.data_as(C.c_void_p), h.value, w.value) == 0
    assert np.array_equal(out.view(np.uint32), ref["gt.pfm"].view(np.uint32))
    q = str(tmp_path / "w.pfm")
    assert lib.smio_write_pfm(q.encode(), out.ctypes.data_as(C.c_void_p), h.value, w.value) == 0
    back = np.empty_like(out)
    assert lib.smio_read_pfm(q.encode(), back.ctypes.data_as(C.c_void_p), h.value, w.value) == 0
    assert np.array_equal(back.view(np.uint32), out.view(np.uint32))


@pytest.mark.parametrize("c", [1, 3])
@pytest.mark.parametrize("fmt", ["png", "pnm"])
def test_write_then_read_is_identity(lib, tmp_path, c, fmt):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (17, 29, 3) if c == 3 else (17, 29), dtype=np.uint8)
    p = str(tmp_path / f"x.{fmt}")
    fn = lib.smio_imwrite_png if fmt == "png" else lib.smio_imwrite_pnm
    assert fn(p.encode(), img.ctypes.data_as(C.c_void_p), 17, 29, c) == 0
    got = _imread(lib, p, 1 if c == 3 else 0)
    assert np.array_equal(got, img)


def test_middlebury_table_and_loader(lib, ref):
    root = os.path.join(IO, "md") + "/"
    h, w, md, found = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib.smio_middlebury_probe(root.encode(), b"Aloe", b".png", C.byref(h), C.byref(w), C.byref(md), C.byref(found)) == 0
    assert md.value == 85                       # maxdispList, main_.cpp:39
    assert found.value == 0b1011                # all, nonocc, DT present; disc mask missing is not fatal (main_.cpp:110-114)
    H, W = h.value, w.value
    I1c, I2c = np.empty((H, W, 3), np.uint8), np.empty((H, W, 3), np.uint8)
    I1g, I2g, al, no = (np.empty((H, W), np.uint8) for _ in range(4))
    DT = np.empty((H, W), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.smio_middlebury_load(root.encode(), b"Aloe", b".png", vp(I1c), vp(I2c), vp(I1g), vp(I2g), vp(al), vp(no), None,
                                    vp(DT)) == 0
    for got, key in ((I1c, "I1_c"), (I2c, "I2_c"), (I1g, "I1_g"), (I2g, "I2_g"), (al, "all"), (no, "nonocc")):
        assert np.array_equal(got, ref["md:" + key]), key
    assert np.array_equal(DT.view(np.uint32), ref["md:DT"].view(np.uint32))
    # unknown object / missing colour images are fatal
    assert lib.smio_middlebury_probe(root.encode(), b"NoSuch", b".png", C.byref(h), C.byref(w), C.byref(md), C.byref(found)) != 0
    assert lib.smio_middlebury_probe(root.encode(), b"Books", b".png", C.byref(h), C.byref(w), C.byref(md), C.byref(found)) != 0


def test_disp_to_bgr_follows_saveDispMap(lib):
    """stereoMatching.h:2005-2110 restated in numpy: min over valid (>= 0) values, max over all, uchar(ratio * (v - min)),
    marker colours, then the error overlay on all_mask pixels with |DT - d| > 1."""
    rng = np.random.default_rng(9)
    H, W = 11, 19
    OCC, MIS, PKR = -2 * 16, -3 * 16, -4 * 16
    d = rng.integers(3, 60, (H, W)).astype(np.int16)
    d[rng.random((H, W)) < 0.1] = OCC
    d[rng.random((H, W)) < 0.1] = MIS
    d[rng.random((H, W)) < 0.05] = PKR
    d[0, 0], d[0, 1], d[0, 2] = -50, -100, -7   # err_ip_dispV, cor_ip_dispV, an unlisted negative (stays black)
    dt = (d + rng.integers(-3, 4, (H, W))).astype(np.float32)
    mask = (rng.random((H, W)) > 0.3).astype(np.uint8) * 255
    valid = d >= 0
    dmin, dmax = d[valid].min(), d.max()
    ratio = np.float32(255.0 / np.float32(dmax - dmin))
    exp = np.zeros((H, W, 3), np.uint8)
    g = (ratio * (d.astype(np.float32) - np.float32(dmin))).astype(np.float32)
    exp[valid] = np.trunc(g[valid]).astype(np.uint8)[:, None]
    exp[d == OCC] = (255, 0, 0)
    exp[d == MIS] = (0, 0, 255)
    exp[d == PKR] = (0, 255, 255)
    exp[d == -50] = (255, 0, 255)
    exp[d == -100] = (255, 255, 0)
    out = np.empty((H, W, 3), np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.smio_disp_to_bgr(vp(d), H, W, OCC, MIS, PKR, vp(out), None, None)
    assert np.array_equal(out, exp)
    exp_err = exp.copy()
    exp_err[(mask > 0) & (np.abs(dt - d.astype(np.float32)) > 1)] = (0, 0, 255)
    lib.smio_disp_to_bgr(vp(d), H, W, OCC, MIS, PKR, vp(out), vp(dt), vp(mask))
    assert np.array_equal(out, exp_err)


@pytest.mark.gpu
def test_middlebury_main_runs_the_reference_driver_flow(lib, tmp_path):
    """tests/cpp/middlebury_main.cpp = the reference's main() (main_.cpp:20-200) over the drop-in class: object folder in,
    saveDispMap picture + calErr line out.  The picture must be the oracle's refined map through saveDispMap's colouring."""
    import subprocess
    from mystereomatching_b200 import synth
    from oracle import pyoracle as po
    exe = os.path.join(ROOT, "mystereomatching_b200", "host", "middlebury_main")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    H, W, D = 64, 96, 60                                   # "teddy": maxdisp 59, ground truth scaled by 4 (main_.cpp:38-39)
    pair = synth.make_pair(H, W, D, "texture_warped", seed=23)
    obj = tmp_path / "md" / "teddy"
    obj.mkdir(parents=True)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    gt8 = np.clip(np.rint(pair["gt"] * 4), 0, 255).astype(np.uint8)
    allm = np.full((H, W), 255, np.uint8)
    for name, a, c in (("im2", pair["bgrL"], 3), ("im6", pair["bgrR"], 3), ("disp2", gt8, 1), ("all", allm, 1),
                       ("nonocc", allm, 1), ("disc", allm, 1)):
        a = np.ascontiguousarray(a)
        assert lib.smio_imwrite_png(str(obj / (name + ".png")).encode(), vp(a), H, W, c) == 0
    raw = tmp_path / "dp.i16"
    r = subprocess.run([exe, str(tmp_path / "md") + "/", "teddy", ".png", "4", str(raw)], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "read-in img done" in r.stdout and "complete teddy" in r.stdout
    # what the class saw: imread(.., 1) and imread(.., 0) of the same colour files
    gL = _imread(lib, str(obj / "im2.png"), 0)
    gR = _imread(lib, str(obj / "im6.png"), 0)
    ref, _, _, _ = po.pipeline(pair["bgrL"], pair["bgrR"], gL, gR, po.default_params(D, paths=4))
    dp = np.fromfile(raw, np.int16).reshape(H, W)
    assert (dp == ref).mean() >= 0.995
    pic = _imread(lib, str(tmp_path / "teddy" / "ADCensus-CBCA-sgm" / "20200627_test_so" / "final.png"), 1)
    exp = np.empty((H, W, 3), np.uint8)
    lib.smio_disp_to_bgr(vp(dp), H, W, -2 * 16, -3 * 16, -4 * 16, vp(exp), None, None)
    assert pic is not None and np.array_equal(pic, exp)
    assert os.path.exists(tmp_path / "teddy" / "ADCensus-CBCA-sgm" / "20200627_test_so" / "final_err.png")
