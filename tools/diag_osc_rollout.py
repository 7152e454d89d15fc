"""Diagnostic: where the time of the OSC action space goes under random actions.  Per policy step:
kernel time of the first simulator step after a NEW action (cold QP partition) and of the nine that follow, rows and QP
iteration statistics.  usage: [CASSIE_ENGINE=quad|thread] python tools/diag_osc_rollout.py [policy_steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200 import envs, lib
n = 16384
P = int(sys.argv[1]) if len(sys.argv) > 1 else 60
env = envs.Cassie2dBatchEnv(n, device=0, task="stand", control_mode="OSC")
env.reset()
lo, hi = env.action_space
g = torch.Generator(device="cuda").manual_seed(0)
lo_t, hi_t = torch.tensor(lo, device="cuda", dtype=torch.float32), torch.tensor(hi, device="cuda", dtype=torch.float32)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def describe(st):
    it = st[:, 2]
    return ("rows mean %.1f max %d | qp iters mean %.2f p99 %.0f max %.0f cta56-max mean %.1f not-optimal %d"
            % (st[:, 0].mean().item(), int(st[:, 0].max().item()), it.mean().item(), torch.quantile(it, 0.99).item(), it.max().item(),
               it[: (n // 56) * 56].reshape(-1, 56).max(dim=1).values.mean().item(), int((st[:, 3] != 0).sum().item())))


tot0 = tot9 = 0.0
for k in range(P):
    # GaussianMLPPolicy at init: mean ~ 0, std 2 on the normalised action, clipped to the box
    raw = torch.randn((n, 7), device="cuda", generator=g) * 2.0
    a = (lo_t + (raw + 1.0) * 0.5 * (hi_t - lo_t)).clamp(lo_t, hi_t)
    ms0 = timed(lambda: env.batch.step_osc(a, 1))
    st0 = env.batch.stats().double()
    ms9 = timed(lambda: env.batch.step_osc(a, 9))
    st9 = env.batch.stats().double()
    tot0 += ms0; tot9 += ms9
    if k % 10 == 0 or k == P - 1:
        z = env.batch.get_general_state()[:, 1]
        print("policy step %3d  first substep %.3f ms [%s]\n                 next nine %.3f ms [%s]  pelvis z mean %.2f below 0.5: %d"
              % (k, ms0, describe(st0), ms9, describe(st9), z.mean().item(), int((z < 0.5).sum().item())))
print("engine %s: %d policy steps: first substeps %.1f ms, the other nine %.1f ms -> %.3g env-steps/s"
      % (os.environ.get("CASSIE_ENGINE", "default"), P, tot0, tot9, n * P * 10 / ((tot0 + tot9) * 1e-3)))
