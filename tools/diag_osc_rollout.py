"""Diagnostic: per-substep QP iteration counts and kernel time of the OSC action space under random actions."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200 import envs, lib
n = 16384
env = envs.Cassie2dBatchEnv(n, device=0, task="stand", control_mode="OSC")
env.reset()
lo, hi = env.action_space
g = torch.Generator(device="cuda").manual_seed(0)
lo_t, hi_t = torch.tensor(lo, device="cuda", dtype=torch.float32), torch.tensor(hi, device="cuda", dtype=torch.float32)
for k in range(6):
    a = lo_t + (hi_t - lo_t) * torch.rand((n, 7), device="cuda", generator=g)
    for s in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.batch.step_osc(a, 1); e1.record(); torch.cuda.synchronize()
        st = env.batch.stats().double()
        it = st[:, 2]
        print("policy step %d substep %d: %.3f ms  rows mean %.1f max %d | qp iters mean %.2f p50 %.0f p99 %.0f max %.0f  warp-max mean %.1f  not-optimal %d"
              % (k, s, e0.elapsed_time(e1), st[:, 0].mean().item(), int(st[:, 0].max().item()), it.mean().item(), it.median().item(),
                 torch.quantile(it, 0.99).item(), it.max().item(), it.reshape(-1, 32).max(dim=1).values.mean().item(), int((st[:, 3] != 0).sum().item())))
