"""Seeded synthetic stereo pairs with known disparity (Middlebury is not available offline).

Two generators, both returning BGR u8 left/right images, the cv2-compatible gray pair, the
left-view ground-truth disparity (float32) and the `all` / `nonocc` masks the reference's
calErr uses (stereoMatching.h:1748-1825):

* random_dot:     right image = i.i.d. uniform u8 BGR; piecewise-constant disparity.
* texture_warped: right image = sum of sinusoids + blurred noise (weak-texture regions, so the
                  CBCA arms actually grow) + per-view noise; one third of the rectangles are
                  slanted planes.

Geometry: background disparity D//8 plus 12 axis-aligned rectangles painted far-to-near; the left
image is the right image forward-warped with a z-buffer, holes filled with fresh texture.
Frame i of a stream uses seed 1000+i.
"""
import numpy as np


def bgr2gray(bgr):
    """cv2.cvtColor(BGR2GRAY) for u8: fixed point with 14 fractional bits."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((1868 * b + 9617 * g + 4899 * r + 8192) >> 14).astype(np.uint8)


def _disparity_layout(rng, H, W, D, slanted):
    d = np.full((H, W), D // 8, np.int32)
    lo, hi = D // 8, max(D // 8 + 1, (3 * D) // 4)
    rects = []
    for k in range(12):
        rw = int(rng.integers(max(2, W // 16), max(3, W // 4) + 1))
        rh = int(rng.integers(max(2, H // 16), max(3, H // 4) + 1))
        x0 = int(rng.integers(0, max(1, W - rw)))
        y0 = int(rng.integers(0, max(1, H - rh)))
        dd = int(rng.integers(lo, hi + 1))
        rects.append((dd, x0, y0, rw, rh, k))
    rects.sort(key=lambda r: r[0])  # far (small d) first, near painted last
    yy, xx = np.mgrid[0:H, 0:W]
    for dd, x0, y0, rw, rh, k in rects:
        sl = (slice(y0, y0 + rh), slice(x0, x0 + rw))
        if slanted and k % 3 == 0:
            a = float(rng.uniform(-0.08, 0.08))
            b = float(rng.uniform(-0.08, 0.08))
            plane = np.rint(dd + a * (xx[sl] - x0 - rw / 2) + b * (yy[sl] - y0 - rh / 2)).astype(np.int32)
            d[sl] = np.clip(plane, lo, D - 1)
        else:
            d[sl] = dd
    return np.clip(d, 0, D - 1)


def _texture(rng, H, W):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.zeros((H, W, 3), np.float32)
    for c in range(3):
        acc = np.full((H, W), 128.0, np.float32)
        for _ in range(6):
            fx, fy = rng.uniform(0.005, 0.08, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += rng.uniform(8, 28) * np.sin(fx * xx + fy * yy + ph).astype(np.float32)
        noise = rng.normal(0, 40, (H, W)).astype(np.float32)
        k = 15
        pad = np.pad(noise, k // 2, mode="edge")
        cs = np.cumsum(np.cumsum(pad, 0), 1)
        cs = np.pad(cs, ((1, 0), (1, 0)))
        box = (cs[k:, k:] - cs[:-k, k:] - cs[k:, :-k] + cs[:-k, :-k]) / (k * k)
        img[..., c] = acc + 6.0 * box[:H, :W]
    return img


def make_pair(H, W, D, kind="random_dot", seed=1000):
    rng = np.random.default_rng(seed)
    slanted = kind == "texture_warped"
    disp = _disparity_layout(rng, H, W, D, slanted)
    if kind == "random_dot":
        right = rng.integers(0, 256, (H, W, 3), dtype=np.uint8).astype(np.float32)
        fill = rng.integers(0, 256, (H, W, 3), dtype=np.uint8).astype(np.float32)
    elif kind == "texture_warped":
        right = _texture(rng, H, W)
        fill = _texture(rng, H, W)
    else:
        raise ValueError(kind)
    # left(v,u) = right(v,u-d(v,u)); out-of-frame sources come from the fill texture
    vv, uu = np.mgrid[0:H, 0:W]
    src = uu - disp
    ok = src >= 0
    left = fill.copy()
    left[ok] = right[vv[ok], src[ok]]
    # visibility: a right pixel is claimed by the nearest (largest d) left pixel mapping to it
    zbuf = np.full((H, W), -1, np.int32)
    np.maximum.at(zbuf, (vv[ok], src[ok]), disp[ok])
    nonocc = np.zeros((H, W), bool)
    nonocc[ok] = zbuf[vv[ok], src[ok]] == disp[ok]
    if kind == "texture_warped":
        left = left + rng.normal(0, 2, left.shape).astype(np.float32)
        right = right + rng.normal(0, 2, right.shape).astype(np.float32)
    left = np.clip(np.rint(left), 0, 255).astype(np.uint8)
    right = np.clip(np.rint(right), 0, 255).astype(np.uint8)
    return {
        "bgrL": np.ascontiguousarray(left), "bgrR": np.ascontiguousarray(right),
        "grayL": bgr2gray(left), "grayR": bgr2gray(right),
        "gt": disp.astype(np.float32), "all": np.ones((H, W), bool), "nonocc": nonocc,
        "H": H, "W": W, "D": D, "kind": kind, "seed": seed,
    }


def bad_k(disp, gt, mask, k=2):
    """Share (in %) of masked pixels with |d - gt| > k or d < 0 (calErr's PBM, stereoMatching.h:1748)."""
    d = disp.astype(np.float32)
    bad = (np.abs(d - gt) > k) | (d < 0)
    m = mask.astype(bool)
    return 100.0 * float(bad[m].sum()) / max(1, int(m.sum()))
