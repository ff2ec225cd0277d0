"""One traced pair of grouped sweeps (SM_B200_LIB = the -DSM_SGMG_TRACE_BUILD variant, SM_SGMG_TRACE = dump path)."""
import sys
sys.path.insert(0, ".")
import torch
from mystereomatching_b200 import capi, synth
H, W, D = 1080, 1920, 256
ctx = capi.Ctx(0)
p = synth.make_pair(H, W, D, "texture_warped", seed=7)
img = ctx.dev(p["bgrL"])
vol = torch.rand((H, W, D), device="cuda")
for _ in range(3):
    ctx.sgm_grouped(vol, img)
torch.cuda.synchronize()
