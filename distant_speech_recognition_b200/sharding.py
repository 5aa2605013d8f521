"""Utterance sharding across the GPUs of one box.

The path shards by independent units: every utterance has its own streams, weights and adaptive state
(reset_stats per utterance, btk20_src/lib/pybeamformer.py:745-762), so rank g simply owns a contiguous utterance range
and there is NO data-path collective.  The only exchange is one all-gather of per-utterance statistics
[sum y^2, frames, NLMS updates] at the end of a run (the reference's "Avg. output power / No. frames processed" report,
unit_test/test_online_beamforming.py:208,336-337, and "Updated weight vectors on n of N frames", pybeamformer.py:751).
torch.distributed is plumbing here (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(total, world, rank):
    """Contiguous [start, stop) utterance range of `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_stats(local_stats, group=None):
    """All-gather per-utterance statistics [U_local][3] (float64) from every rank, in rank order -> [U_total][3].
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local_stats, np.float64).reshape(-1, 3)
    if not (dist.is_available() and dist.is_initialized()):
        return local.copy()
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts) if counts else 0
    buf = torch.zeros((cap, 3), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def summarize(stats, frame_shift):
    """The reference's end-of-run report from gathered statistics."""
    stats = np.asarray(stats, np.float64).reshape(-1, 3)
    frames = float(stats[:, 1].sum())
    return {"utterances": int(stats.shape[0]), "frames": frames,
            "avg_output_power": float(stats[:, 0].sum() / max(frames * frame_shift, 1.0)),
            "nlms_updates": float(stats[:, 2].sum())}
