"""BASELINE.json configs[4] (TRPO rollout collection): fused policy MLP + noise + env step kernel, then the
sampler-side pre-processing (returns, LinearFeatureBaseline fit, GAE advantages) on device.  Not the headline
bench (bench.py); prints one JSON line for profiles/.   usage: python tools/bench_rollout.py [--envs N] [--T T]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200.rollout import GaussianMLPPolicy, RolloutCollector  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--envs", type=int, default=16384)
p.add_argument("--T", type=int, default=20, help="policy steps per collect() (10 sim steps each)")
p.add_argument("--reps", type=int, default=5)
p.add_argument("--mode", default="PD", choices=["PD", "OSC", "TORQUE"])
p.add_argument("--task", default="stand", choices=["stand", "imitate"])
a = p.parse_args()

col = RolloutCollector(a.envs, device=0, task=a.task, control_mode=a.mode, max_path_length=1000)
pol = GaussianMLPPolicy(col.obs_dim, col.act_dim)
col.collect(pol, a.T)                      # warm-up (also sizes the buffers)
ret = col.discounted_returns()
col.fit_baseline(ret); col.advantages()
torch.cuda.synchronize()


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(a.reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


ms_collect = timed(lambda: col.collect(pol, a.T))
t0 = time.perf_counter()
for _ in range(a.reps):
    ret = col.discounted_returns(); col.fit_baseline(ret); col.advantages()
torch.cuda.synchronize()
ms_post = (time.perf_counter() - t0) / a.reps * 1e3
d = col.done
st = col.batch.stats().double()   # [n, 4]: rows, PGS sweeps, QP iterations, QP status of the last sim step
qp = {"iters_mean": float(st[:, 2].mean().item()), "iters_max": int(st[:, 2].max().item()),
      "iters_p99": float(torch.quantile(st[:, 2], 0.99).item()), "not_optimal": int((st[:, 3] != 0).sum().item()),
      "warp_max_mean": float(st[:, 2].reshape(-1, 32).max(dim=1).values.mean().item())}
print(json.dumps({
    "workload": "rollout: GaussianMLPPolicy(%d->32->32->%d) + %s action space + cassie2d %s env, %d envs x %d policy steps x 10 sim steps"
                % (col.obs_dim, col.act_dim, a.mode, a.task, a.envs, a.T),
    "collect_ms": ms_collect, "env_steps_per_s": a.envs * a.T * 10 / (ms_collect * 1e-3),
    "policy_steps_per_s": a.envs * a.T / (ms_collect * 1e-3),
    "returns_baseline_advantages_ms": ms_post, "episodes_finished": int((d != 0).sum().item()),
    "last_step_rows_mean": float(st[:, 0].mean().item()), "last_step_qp": qp}))
