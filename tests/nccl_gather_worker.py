"""Worker of tests/test_gpu_nccl.py (launched by torch.distributed.run, one rank per GPU): shards a rollout by global env
id, gathers the sample paths and reduces the statistics over NCCL, and checks both against what a single GPU collects
for the same global env ids (Philox streams are keyed by global env id, so the result must not depend on the GPU count)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200 import parallel  # noqa: E402
from cassierl_b200.rollout import GaussianMLPPolicy, RolloutCollector  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_total, T = 96 + 1, 12                       # ragged: shards differ by one env
a, b = parallel.shard_range(n_total, rank, world)
col = RolloutCollector(b - a, device=local, control_mode="PD", precision=64, first_global_env=a)
pol = GaussianMLPPolicy(col.obs_dim, col.act_dim, device=dev, dtype=torch.float64)
dist.broadcast(pol.flat, src=0)
col.collect(pol, T)
stats = parallel.RolloutStats(device=dev)
stats.update(col.rew, col.done, col.obs)
red = stats.reduce()
g_obs = parallel.gather_paths(col.obs, n_total)
g_rew = parallel.gather_paths(col.rew, n_total)
g_done = parallel.gather_paths(col.done, n_total)
assert g_obs.shape == (T, n_total, col.obs_dim) and g_done.dtype == torch.uint8
if rank == 0:
    ref = RolloutCollector(n_total, device=local, control_mode="PD", precision=64, first_global_env=0)
    ref.collect(pol, T)
    assert torch.equal(ref.done, g_done), "done flags differ between 1 GPU and %d GPUs" % world
    assert (ref.obs - g_obs).abs().max().item() < 1e-12 and (ref.rew - g_rew).abs().max().item() < 1e-12
    assert abs(red["reward_sum"] - ref.rew.double().sum().item()) < 1e-9 * max(1.0, abs(red["reward_sum"]))
    assert red["steps"] == T * n_total and red["non_finite"] == 0
    print("NCCL_GATHER_OK world=%d envs=%d" % (world, n_total), flush=True)
dist.barrier()
dist.destroy_process_group()
