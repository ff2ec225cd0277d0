"""Times sm_region_vote (unanimity path vs histogram path) on a 1080p LR-checked disparity map."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mystereomatching_b200 import capi, synth
H, W, D = 1080, 1920, 256
p = synth.make_pair(H, W, D, "texture_warped", seed=1000)
ctx = capi.Ctx(0)
params = capi.default_params(D - 1, sgm_paths=8, Do_regionVote=0, Do_properIpol=0, Do_lastMedianBlur=0)
pl = capi.Pipeline(ctx, H, W, params)
d = pl.run(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"]).copy()     # LR-checked, not voted
print("invalid fraction", float((d < 0).mean()))
arms = pl.buffer(4, (H, W, 5), torch.int16).clone()
for ratio in (0.4, 0.0):
    dd = ctx.dev(d.copy())
    for _ in range(3): ctx.region_vote(dd.clone(), arms, D, ratio, 20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ins = [dd.clone() for _ in range(20)]
    torch.cuda.synchronize(); e0.record()
    for x in ins: ctx.region_vote(x, arms, D, ratio, 20)
    e1.record(); torch.cuda.synchronize()
    print("ratio", ratio, "ms per call", e0.elapsed_time(e1) / 20)
