"""Session-parallel sharding (SURVEY.md §8 e): client sessions / renderer streams are independent,
so scaling across the GPUs of a box is a partition of whole sessions -- session s runs on GPU
s % n_gpus -- with NO data-path collective (nothing is reduced across GPUs).  torch.distributed is
only used by the benchmark harness for the start barrier and the max-over-ranks of the timings.
"""
from __future__ import annotations


def device_for_session(session_id: int, n_gpus: int) -> int:
    """GPU that owns a client session (the reference runs one process per session, main.cpp:133-171;
    here one process per GPU serves every session of its shard)."""
    if n_gpus < 1:
        raise ValueError("n_gpus must be >= 1")
    return session_id % n_gpus


def sessions_of_rank(rank: int, world: int, n_sessions: int) -> list:
    """The shard of rank `rank`: every session whose device is `rank`."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return [s for s in range(n_sessions) if device_for_session(s, world) == rank]


def aggregate_fps(frames_per_rank: list, seconds_per_rank: list) -> float:
    """Whole-job throughput: all frames of all ranks over the slowest rank's time."""
    return float(sum(frames_per_rank)) / max(seconds_per_rank)


def max_over_ranks(value: float, dist=None) -> float:
    """MAX all-reduce of a timing (gloo on CPU, nccl on GPU); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return value
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
