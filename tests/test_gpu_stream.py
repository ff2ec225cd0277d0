"""The frame stream of the C ABI (sm_stream_*, SURVEY.md 8e / BASELINE config 5): frames submitted in order, frame i on
device i mod n, uploads and downloads overlapped with the neighbours' compute -- every map must equal the one the plain
host-buffer call (sm_pipeline_run) gives for the same pair, whatever the order of waiting, the buffer kind (pinned /
pageable) and the number of devices."""
import os
import subprocess

import numpy as np
import pytest
import torch

from mystereomatching_b200 import capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(n, H, W, D):
    return [synth.make_pair(H, W, D, "texture_warped", seed=1000 + i) for i in range(n)]


def _reference_maps(ctx, frames, H, W, D, **over):
    pl = capi.Pipeline(ctx, H, W, capi.default_params(D - 1, **over))
    maps = [pl.run(f["bgrL"], f["bgrR"], f["grayL"], f["grayR"]).copy() for f in frames]
    pl.close()
    return maps


@pytest.mark.timeout(600)
@pytest.mark.parametrize("ndev", [1, 2, 8])
def test_stream_equals_frame_by_frame(ctx, ndev):
    have = torch.cuda.device_count()
    if ndev > have:
        pytest.skip(f"{have} GPU(s) visible")
    H, W, D, n = 96, 160, 160, 11          # D = 160: the grouped sweeps (cooperative launch) are on the path
    frames = _frames(n, H, W, D)
    want = _reference_maps(ctx, frames, H, W, D, sgm_paths=8)
    st = capi.Stream(list(range(ndev)), H, W, capi.default_params(D - 1, sgm_paths=8), queue_depth=3)
    outs, tickets = [], []
    for i, f in enumerate(frames):
        pinned = i % 2 == 0                 # alternate pinned and pageable buffers
        arrs = [torch.from_numpy(f[k]).pin_memory().numpy() if pinned else f[k] for k in ("bgrL", "bgrR", "grayL", "grayR")]
        if i % 3 == 2:
            arrs[2] = arrs[3] = None        # gray computed on the device from BGR (cv2-compatible: same bits)
        out = torch.empty((H, W), dtype=torch.int16).pin_memory().numpy() if pinned else np.empty((H, W), np.int16)
        outs.append(out)
        tickets.append(st.submit(arrs[0], arrs[1], arrs[2], arrs[3], out))
    for t in reversed(tickets):             # wait in reverse order
        st.wait(t)
    st.drain()
    done = [st.frames_done(k) for k in range(ndev)]
    assert sum(done) == n and max(done) - min(done) <= 1, done      # frame i -> worker i mod n
    assert st.launches() > 0
    st.close()
    for i in range(n):
        assert np.array_equal(outs[i], want[i]), i


@pytest.mark.timeout(600)
def test_stream_from_plain_c(tmp_path):
    exe = os.path.join(ROOT, "mystereomatching_b200", "host", "stream_main")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    H, W, D, n = 64, 96, 32, 7
    frames = _frames(n, H, W, D)
    with open(tmp_path / "in.bin", "wb") as f:
        for fr in frames:
            for k in ("bgrL", "bgrR", "grayL", "grayR"):
                f.write(np.ascontiguousarray(fr[k]).tobytes())
    ndev = min(2, torch.cuda.device_count())
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(H), str(W), str(D), "4", str(n), str(ndev)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-500:]
    assert f"frames {n} " in r.stdout
    got = np.fromfile(tmp_path / "out.bin", np.int16).reshape(n, H, W)
    c = capi.Ctx(0)
    want = _reference_maps(c, frames, H, W, D, sgm_paths=4)
    c.close()
    for i in range(n):
        assert np.array_equal(got[i], want[i]), i


def test_stream_argument_errors():
    p = capi.default_params(31)
    with pytest.raises(capi.SmError):
        capi.Stream([99], 32, 32, p)
    with pytest.raises(capi.SmError):
        capi.Stream([], 32, 32, p)
