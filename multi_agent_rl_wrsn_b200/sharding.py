"""Sharding environments over the GPUs of one box.

The reference is single-process; here environments are independent, so a job of ``num_envs``
environments splits into contiguous blocks of environment indices, one process per GPU, with NO collective on the
simulation path.  The only cross-rank traffic is the rollout statistics (and, in a trainer, the gradient all-reduce):
``reduce_stats`` is a plain ``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(num_envs, rank, world_size):
    """[lo, hi) of the environment indices owned by ``rank``: contiguous blocks, sizes differing by at most one."""
    base, rem = divmod(int(num_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_scenario_index(num_envs, num_scenarios, rank, world_size):
    """Scenario index of every local environment such that the global job is identical for every world size:
    global environment e always uses scenario e % num_scenarios."""
    lo, hi = shard_range(num_envs, rank, world_size)
    return (np.arange(lo, hi) % int(num_scenarios)).astype(np.int32)


def reduce_stats(env, group=None):
    """Job-wide (decisions, simulated seconds): sum of the kernels' running totals over environments and ranks
    (the third total, resets, is ``env.req.stats[:, 2]``)."""
    s = env.req.stats.sum(0).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return float(s[0].item()), float(s[1].item())
