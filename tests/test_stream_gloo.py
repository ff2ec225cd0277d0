"""N>1 host logic on CPU: frame -> rank partitioning, max-over-ranks timing and result gathering, with two gloo
processes.  (The data path itself has no collective; see mystereomatching_b200/stream.py.)"""
import os
import subprocess
import sys
import textwrap

from mystereomatching_b200 import stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round_robin_partition_covers_every_frame_once():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            seen = []
            for r in range(world):
                fr = stream.frames_of_rank(n, r, world)
                assert all(stream.owner_of_frame(i, world) == r for i in fr)
                seen += fr
            assert sorted(seen) == list(range(n))
            sizes = [len(stream.frames_of_rank(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_stream():
    out, dt = stream.run_stream(lambda i: i * i, 5)
    assert out == {i: i * i for i in range(5)} and dt >= 0


WORKER = textwrap.dedent("""
    import os, sys, time
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    from mystereomatching_b200 import stream, synth
    from oracle import pyoracle as po
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    def process(i):            # a frame of the stream: seed 1000 + i (SURVEY.md 8d), CPU oracle as the stand-in worker
        p = synth.make_pair(24, 40, 8, "random_dot", seed=1000 + i)
        if rank == 1:
            time.sleep(0.05)   # rank 1 is slower: the reported time must be ITS time
        d, _, _, _ = po.pipeline(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], po.default_params(8))
        return int(d.astype(np.int64).sum())
    out, secs = stream.run_stream(process, 5, dist)
    if rank == 0:
        ref = {{}}
        for i in range(5):
            p = synth.make_pair(24, 40, 8, "random_dot", seed=1000 + i)
            d, _, _, _ = po.pipeline(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], po.default_params(8))
            ref[i] = int(d.astype(np.int64).sum())
        assert out == ref, (out, ref)
        assert secs >= 0.1, secs     # 2 frames x 0.05 s on the slow rank
        print("STREAM_OK", secs)
    else:
        assert out is None
    dist.destroy_process_group()
""")


def test_two_rank_stream_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True,
                       text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "STREAM_OK" in r.stdout
