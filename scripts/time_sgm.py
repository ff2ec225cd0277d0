"""Time sm_sgm (per-path kernels) against sm_sgm_grouped on one synthetic volume.  usage: time_sgm.py [H W D]"""
import sys, time
sys.path.insert(0, ".")
import torch
from mystereomatching_b200 import capi, synth
H, W, D = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1080, 1920, 256)
ctx = capi.Ctx(0)
p = synth.make_pair(H, W, D, "texture_warped", seed=7)
img = ctx.dev(p["bgrL"])
vol = torch.rand((H, W, D), device="cuda")
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
img2 = ctx.dev(p["bgrR"])
vol2 = torch.rand((H, W, D), device="cuda")
t2 = timeit(lambda: ctx.sgm_grouped2(vol, vol2, img, img2))
print(f"{W}x{H} D={D}: sm_sgm_grouped2 (two views) {t2:.3f} ms = {t2 / 2:.3f} per view")
print(f"{W}x{H} D={D}: sm_sgm(8) {timeit(lambda: ctx.sgm(vol, img, 8)):.3f} ms   sm_sgm_grouped {timeit(lambda: ctx.sgm_grouped(vol, img)):.3f} ms")
