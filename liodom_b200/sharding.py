"""Multi-GPU plumbing of the hot path: independent sequences are sharded across ranks (one
process per GPU), there is no data-path collective; torch.distributed carries only the timing
barrier, the max-over-ranks of the device time and the final gather of poses."""
import numpy as np


def shard_sequences(n_sequences, world, rank):
    """Contiguous block partition of sequence ids 0..n-1; the first n % world ranks get one more."""
    base, extra = divmod(n_sequences, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def seed_of(sequence_id, first_seed=1000):
    """Synthetic sequences use seeds 1000, 1001, ...."""
    return first_seed + sequence_id


def max_over_ranks(value, group=None):
    """Job time = slowest rank (device-measured milliseconds)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def job_throughput(units_per_rank, ms_local, world, group=None):
    """Whole-job units/s: all ranks' units divided by the slowest rank's time."""
    ms = max_over_ranks(ms_local, group)
    return world * units_per_rank / (ms * 1e-3), ms


def gather_poses(local_poses, local_ids, n_sequences, group=None):
    """All ranks' trajectories on every rank, indexed by sequence id: [n_sequences, frames, 4, 4]."""
    import torch
    import torch.distributed as dist
    local_poses = np.asarray(local_poses, np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = np.zeros((n_sequences,) + local_poses.shape[1:])
        out[local_ids] = local_poses
        return out
    objs = [None] * dist.get_world_size(group)
    dist.all_gather_object(objs, (list(local_ids), local_poses), group=group)
    out = np.zeros((n_sequences,) + local_poses.shape[1:])
    for ids, poses in objs:
        if len(ids):
            out[ids] = poses
    return out
