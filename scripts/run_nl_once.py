import sys
sys.path.insert(0, ".")
import torch
from mystereomatching_b200 import capi, synth
ctx = capi.Ctx(0)
H, W, D = 480, 640, 64
p = synth.make_pair(H, W, D, "texture_warped", seed=1000)
img = ctx.dev(p["bgrL"])
vol = torch.rand((H, W, D), device="cuda")
for _ in range(2):
    ctx.nl(img, vol)
torch.cuda.synchronize()
