"""Which constraint tier do the envs of an OSC-action rollout sit in?  (quad_engine.cuh: tier 0 / 1 / 2 / serial fallback)
After every collect() the joint-limit count per leg is recomputed from qpos and the contact count follows from the
row count of the last simulator step (rows = 4 + limits + 2 contacts).  Prints a histogram per collect.

  python tools/diag_tiers.py [--envs N] [--collects K]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200.rollout import GaussianMLPPolicy, RolloutCollector  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--envs", type=int, default=16384)
p.add_argument("--collects", type=int, default=10)
p.add_argument("--T", type=int, default=20)
a = p.parse_args()
LO = np.radians([-50, -164, 50, -140]); HI = np.radians([80, -37, 170, -30])   # hip, knee, tarsus, toe (cassie2d_stiff.xml:77-95)
col = RolloutCollector(a.envs, device=0, task="stand", control_mode="OSC", max_path_length=1000)
pol = GaussianMLPPolicy(col.obs_dim, col.act_dim, device=torch.device("cuda", 0))
for c in range(a.collects):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); col.collect(pol, a.T); e1.record(); torch.cuda.synchronize()
    g = col.batch.get_general_state().double().cpu().numpy()   # [x z pitch | their rates | left q(5) | left qd(5) | right q(5) | right qd(5)]
    q = np.concatenate([g[:, 0:3], g[:, 6:11], g[:, 16:21]], axis=1)
    rows = col.batch.stats().double().cpu().numpy()[:, 0]
    nl = []
    for L in range(2):
        ql = q[:, 3 + 5 * L: 3 + 5 * L + 4]
        nl.append(((ql < LO) | (ql > HI)).sum(axis=1))
    nlmax = np.maximum(nl[0], nl[1]); nlim = nl[0] + nl[1]
    ncon = np.round((rows - 4 - nlim) / 2)
    print("collect %2d  %.1f ms (%.3g env-steps/s)  limits on the worse leg: %s   contacts (both legs): %s   pelvis z mean %.2f"
          % (c, e0.elapsed_time(e1), a.envs * a.T * 10 / (e0.elapsed_time(e1) * 1e-3),
             {int(k): int((nlmax == k).sum()) for k in np.unique(nlmax)},
             {int(k): int((ncon == k).sum()) for k in np.unique(ncon)}, q[:, 1].mean()), flush=True)
