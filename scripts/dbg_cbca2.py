import numpy as np, torch, sys
sys.path.insert(0, ".")
from mystereomatching_b200 import capi
from oracle import pyoracle as po
ctx = capi.Ctx(0)
rng = np.random.default_rng(0)
for (H, W, D) in [(5, 20, 4), (37, 53, 19), (64, 160, 64), (90, 200, 40)]:
    bL = rng.integers(90, 120, (H, W, 3), dtype=np.uint8); bR = rng.integers(90, 120, (H, W, 3), dtype=np.uint8)
    bL[:, : W // 2] = 100; bR[:, : W // 2] = 100
    aL, aR = po.arms(bL), po.arms(bR)
    vol = rng.random((H, W, D)).astype(np.float32)
    for view in (0, 1):
        for iters in (1, 2):
            ref = po.cbca(vol, aL, aR, iters, view)
            got = ctx.cbca(ctx.dev(vol.copy()), ctx.dev(aL.view(np.int16)), ctx.dev(aR.view(np.int16)), iters, view).cpu().numpy()
            bad = np.argwhere(got.view(np.uint32) != ref.view(np.uint32))
            print((H, W, D), "view", view, "iters", iters, "mismatch", len(bad), "of", ref.size, bad[:4].tolist(), flush=True)
