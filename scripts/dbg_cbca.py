import numpy as np, torch, sys
sys.path.insert(0, ".")
from mystereomatching_b200 import capi, synth
from oracle import pyoracle as po
ctx = capi.Ctx(0)
rng = np.random.default_rng(0)
for (H, W, D) in [(1, 53, 19), (37, 1, 19), (1, 53, 32), (5, 20, 4), (37, 53, 19), (37, 53, 32)]:
    for flat in (True, False):
        if flat:
            bL = np.full((H, W, 3), 100, np.uint8); bR = bL.copy()
        else:
            bL = rng.integers(90, 120, (H, W, 3), dtype=np.uint8); bR = rng.integers(90, 120, (H, W, 3), dtype=np.uint8)
        aL, aR = po.arms(bL), po.arms(bR)
        vol = rng.integers(0, 8, (H, W, D)).astype(np.float32)
        for view in (0, 1):
            ref = po.cbca(vol, aL, aR, 1, view)
            got = ctx.cbca(ctx.dev(vol.copy()), ctx.dev(aL.view(np.int16)), ctx.dev(aR.view(np.int16)), 1, view).cpu().numpy()
            bad = np.argwhere(got != ref)
            print((H, W, D), "flat" if flat else "rand", "view", view, "mismatch", len(bad), "of", ref.size, bad[:5].tolist(),
                  [(float(got[tuple(b)]), float(ref[tuple(b)])) for b in bad[:3]])
