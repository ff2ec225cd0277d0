"""Frame-parallel streaming over the GPUs of one box.

Stereo frames are independent units; inside a frame the SGM paths span the image in both axes and the MST is
global, so the path does not shard below a frame and has no exchange step.  Partitioning is therefore by frame,
round-robin `frame i -> rank i mod N`, one process per GPU, no data-path collective.  torch.distributed is used
only for what a driver needs around the data path: a barrier, the max-over-ranks time and gathering the (small)
per-frame results on rank 0.

The functions take a `process(frame_index) -> result` callable so the host logic can be tested on CPU with the
gloo backend (tests/test_stream_gloo.py); on the GPU box the callable wraps capi.Pipeline.run.
"""
import time


def frames_of_rank(n_frames, rank, world):
    """Indices of the frames rank `rank` processes: i with i mod world == rank, in stream order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def owner_of_frame(i, world):
    return i % world


def run_stream(process, n_frames, dist=None, device=None):
    """Run `process(i)` for this rank's frames; returns (results_by_frame on rank 0 else None, seconds = max over ranks).

    dist: the torch.distributed module of an initialised process group, or None for a single process."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    mine = frames_of_rank(n_frames, rank, world)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    local = [(i, process(i)) for i in mine]
    dt = time.perf_counter() - t0
    if dist is None:
        return dict(local), dt
    import torch
    t = torch.tensor([dt], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None, float(t.item())
    out = {}
    for part in gathered:
        for i, r in part:
            if i in out:
                raise RuntimeError(f"frame {i} processed twice")
            out[i] = r
    if sorted(out) != list(range(n_frames)):
        raise RuntimeError("frames missing from the stream")
    return out, float(t.item())
