import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from mystereomatching_b200 import capi, synth
ctx = capi.Ctx(0)
torch.manual_seed(0)
CONFIGS = [(480, 640, 64, "texture_warped"), (480, 640, 64, "random_dot"), (1080, 1920, 64, "texture_warped")]
if len(sys.argv) > 1: CONFIGS = CONFIGS[:int(sys.argv[1])]
for (H, W, D, kind) in CONFIGS:
    p = synth.make_pair(H, W, D, kind, seed=1000)
    img = ctx.dev(p["bgrL"])
    vol = torch.rand((H, W, D), device="cuda")
    def timeit(f, n=3):
        f(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n): f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    l0 = ctx.launches()
    tree = ctx.mst_build(img)
    torch.cuda.synchronize()
    nl_launch = ctx.launches() - l0
    t_mst = timeit(lambda: ctx.mst_build(img))
    t_tf = timeit(lambda: ctx.tree_filter(vol, tree, 0.1))
    t_nl = timeit(lambda: ctx.nl(img, vol))
    depth = int(tree["rank"].max().item()) + 1
    import hashlib
    sha = hashlib.sha1(ctx.tree_filter(vol, tree, 0.1).cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"{W}x{H} D={D} {kind}: depth {depth}  mst_build {t_mst:.2f} ms ({nl_launch} launches)  tree_filter {t_tf:.2f} ms  sm_nl {t_nl:.2f} ms  sha {sha}", flush=True)
