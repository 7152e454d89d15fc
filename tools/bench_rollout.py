"""BASELINE.json configs[4] (TRPO rollout collection): fused policy MLP + noise + env step kernel, then the
sampler-side pre-processing (returns, LinearFeatureBaseline fit, GAE advantages) on device.  Not the headline
bench (bench.py); prints one JSON line for profiles/.

  python tools/bench_rollout.py [--envs N] [--T T] [--mode PD|OSC|TORQUE]
  python -m torch.distributed.run --nproc-per-node G ... tools/bench_rollout.py     # N envs PER GPU, one rank per GPU

Multi-GPU: envs are sharded by global env id (Philox streams and phases do not depend on G), there is no
collective inside a rollout; after it, NCCL all-reduces the rollout statistics (cassierl_b200.parallel.RolloutStats) and
all-gathers the sample paths for a single learner (parallel.gather_paths) on a side stream, overlapped with the next
rollout.  Times are CUDA-event times, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200 import parallel  # noqa: E402
from cassierl_b200.rollout import GaussianMLPPolicy, RolloutCollector  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--envs", type=int, default=16384, help="envs per GPU")
p.add_argument("--T", type=int, default=20, help="policy steps per collect() (10 sim steps each)")
p.add_argument("--reps", type=int, default=5)
p.add_argument("--mode", default="PD", choices=["PD", "OSC", "TORQUE"])
p.add_argument("--task", default="stand", choices=["stand", "imitate"])
a = p.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_total = a.envs * world

col = RolloutCollector(a.envs, device=local, task=a.task, control_mode=a.mode, max_path_length=1000, first_global_env=rank * a.envs)
pol = GaussianMLPPolicy(col.obs_dim, col.act_dim, device=dev)
if world > 1:
    dist.broadcast(pol.flat, src=0)            # the learner's parameters (8.5 kB), once per iteration
col.collect(pol, a.T)                      # warm-up (also sizes the buffers)
ret = col.discounted_returns()
col.fit_baseline(ret); col.advantages()
torch.cuda.synchronize()


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return maxr(sorted(ms)[len(ms) // 2])


ms_collect = timed(lambda: col.collect(pol, a.T), a.reps)
t0 = time.perf_counter()
for _ in range(a.reps):
    ret = col.discounted_returns(); col.fit_baseline(ret); col.advantages()
torch.cuda.synchronize()
ms_post = maxr((time.perf_counter() - t0) / a.reps * 1e3)

# ---- what a single learner needs from all ranks: statistics (all-reduce) and the sample paths (all-gather)
stats = parallel.RolloutStats(device=dev)
stats.update(col.rew, col.done, col.obs)
red = stats.reduce()
gather = {"world": world, "collective": "none (1 GPU)"}
if world > 1:
    fields = [col.obs, col.act, col.mean, col.rew, col.done]
    local_bytes = sum(x.numel() * x.element_size() for x in fields)

    def do_gather():
        return [parallel.gather_paths(x, n_total) for x in fields]

    g = do_gather()
    assert g[0].shape == (a.T, n_total, col.obs_dim)
    # rank r's shard sits at its global env ids
    assert torch.equal(g[3][:, rank * a.envs:(rank + 1) * a.envs], col.rew)
    ms_gather = timed(do_gather, a.reps)
    # overlapped: the gather of batch k runs on a side stream while batch k+1 is being collected
    side = torch.cuda.Stream(device=dev)
    snap = [x.clone() for x in fields]

    def overlapped():
        for i, x in enumerate(fields):
            snap[i].copy_(x)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            out = [parallel.gather_paths(x, n_total) for x in snap]
        col.collect(pol, a.T)
        torch.cuda.current_stream().wait_stream(side)
        return out

    ms_overlap = timed(overlapped, a.reps)
    gather = {"world": world, "collective": "ncclAllGather (torch.distributed all_gather) of obs/act/mean/rew/done + ncclAllReduce of 6 statistics",
              "bytes_per_rank": local_bytes, "bytes_gathered_per_rank": local_bytes * world, "gather_ms": ms_gather,
              "gather_GBps_per_rank_in": local_bytes * (world - 1) / (ms_gather * 1e-3) / 1e9,
              "collect_plus_gather_serial_ms": ms_collect + ms_gather, "collect_with_overlapped_gather_ms": ms_overlap}

d = col.done
st = col.batch.stats().double()   # [n, 4]: rows, PGS sweeps, QP iterations, QP status of the last sim step
qp = {"iters_mean": float(st[:, 2].mean().item()), "iters_max": int(st[:, 2].max().item()),
      "iters_p99": float(torch.quantile(st[:, 2], 0.99).item()), "not_optimal": int((st[:, 3] != 0).sum().item()),
      "warp_max_mean": float(st[:, 2].reshape(-1, 32).max(dim=1).values.mean().item())}
if rank == 0:
    print(json.dumps({
        "workload": "rollout: GaussianMLPPolicy(%d->32->32->%d) + %s action space + cassie2d %s env, %d envs per GPU x %d GPUs x %d policy steps x 10 sim steps"
                    % (col.obs_dim, col.act_dim, a.mode, a.task, a.envs, world, a.T),
        "mlp": os.environ.get("CASSIE_MLP", "scalar"), "engine": os.environ.get("CASSIE_ENGINE", "default"),
        "n_gpus": world, "collect_ms": ms_collect, "env_steps_per_s": n_total * a.T * 10 / (ms_collect * 1e-3),
        "policy_steps_per_s": n_total * a.T / (ms_collect * 1e-3),
        "env_steps_per_s_with_overlapped_gather": (n_total * a.T * 10 / (gather["collect_with_overlapped_gather_ms"] * 1e-3)) if world > 1 else None,
        "returns_baseline_advantages_ms": ms_post, "episodes_finished": int((d != 0).sum().item()),
        "rollout_stats_all_ranks": red, "gather": gather,
        "last_step_rows_mean": float(st[:, 0].mean().item()), "last_step_qp": qp}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
