"""Stream-level sharding for multi-GPU runs: frames of one stream form a
dependency chain and a frame's wavefront does not split across devices, so GPUs only ever
take whole, independent streams.  No data-path collective exists; torch.distributed is used
for the start barrier and to combine the per-rank clocks and frame counts."""
import os


def rank_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def streams_for_rank(rank, streams_per_gpu, n_unique):
    """Weak scaling: every rank decodes `streams_per_gpu` independent decoder instances.
    Instance i of rank r plays clip (r*streams_per_gpu + i) mod n_unique, so different ranks
    start at different clips when fewer unique clips than instances exist."""
    base = rank * streams_per_gpu
    return [(base + i) % n_unique for i in range(streams_per_gpu)]


def combine(dist, device, frames, seconds):
    """Whole-job figures: total frames over all ranks, and the MAX of the per-rank times."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return frames, seconds
    import torch
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    f = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return float(f.item()), float(t.item())
