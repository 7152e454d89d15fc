"""Env-batch data parallelism: one process per GPU, envs sharded contiguously, no collective on the
step path.  The reference samples with one worker (n_parallel=1: rllab/envs/trpo_cassie.py:48); the
only cross-rank traffic here is what a single learner needs after a rollout: reduced statistics and,
optionally, the gathered sample paths.  Works on any torch.distributed backend (NCCL on
the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous [start, stop) of global env ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_env_ids(n_total, rank=None, world=None, device="cpu"):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    a, b = shard_range(n_total, rank, world)
    return torch.arange(a, b, device=device)


def squat_phases(ids, n_total, dtype=torch.float32):
    """Per-env phase offset of the squatting target, keyed by GLOBAL env id (BASELINE configs[2]):
    phi_e = 2 pi e / N, so results are independent of how many GPUs share the batch."""
    return (2.0 * torch.pi * ids.to(torch.float64) / float(n_total)).to(dtype)


class RolloutStats:
    """Running sums a learner wants from a rollout: reward, path length, episodes, non-finite envs."""

    FIELDS = ("reward_sum", "steps", "episodes", "non_finite", "terminated", "truncated")

    def __init__(self, device="cpu"):
        self.acc = torch.zeros(len(self.FIELDS), dtype=torch.float64, device=device)

    def update(self, reward, done, obs=None):
        """done: bool, or the rollout kernel's uint8 flag (1 = env terminated, 2 = max_path_length reached)"""
        self.acc[0] += reward.double().sum()
        self.acc[1] += reward.numel()
        self.acc[2] += (done != 0).double().sum()
        self.acc[4] += (done == 1).double().sum()
        self.acc[5] += (done == 2).double().sum()
        if obs is not None:
            self.acc[3] += (~torch.isfinite(obs).all(dim=-1)).double().sum()

    def reduce(self, group=None):
        """All-reduce (sum) over ranks; returns a dict of Python floats, identical on every rank."""
        t = self.acc.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        d = dict(zip(self.FIELDS, t.tolist()))
        d["mean_reward"] = d["reward_sum"] / max(d["steps"], 1.0)
        d["mean_path_length"] = d["steps"] / max(d["episodes"], 1.0)
        return d


def gather_paths(local, n_total, group=None):
    """All-gather a per-env tensor [T, n_local, ...] into global env order [T, n_total, ...].
    Shards may differ in size by one env, so ranks pad to the largest shard and trim after."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    pad = local
    if local.shape[1] < nmax:
        fill = torch.zeros((local.shape[0], nmax - local.shape[1]) + tuple(local.shape[2:]), dtype=local.dtype, device=local.device)
        pad = torch.cat([local, fill], dim=1)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:, : b - a] for o, (a, b) in zip(out, sizes)], dim=1)
