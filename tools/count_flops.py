"""Algorithmic FLOPs of one simulator step per control mode (bench.py FLOPS_PER_STEP).

The device engine (cassierl_b200/csrc/*.cuh) is instantiated on an operation-counting scalar by the
test harness (tests/host_harness, hh_count_ops) and run along the squatting stream; add/sub, mul, div
and sqrt count 1 each (an FMA therefore 2), transcendental calls are listed separately.  The OSC
QP runs in plain double and is added analytically: G/g assembly 105*11*3 + 14*11*3 flops, and per
block-pivoting iteration one 14x14 Cholesky (14^3/3), two triangular solves and two mat-vecs.
"""
import ctypes as ct
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Harness, squat_jacobian_action, squat_osc_action, QPOS_INIT_PY  # noqa: E402
from oracle import oracle as O  # noqa: E402

QP_ASSEMBLY = 105 * 11 * 3 + 14 * 11 * 3
QP_ITER = 14 ** 3 // 3 + 2 * 14 * 14 + 2 * 2 * 14 * 14


def main(n_steps=100):
    h = Harness(os.path.join(ROOT, "tests", "_build", "libhost_harness.so"), O.default_model_path())
    m = O.Model()
    lp = ct.POINTER(ct.c_long)
    # count the algorithm, not the fast path's inert padding rows (those are wasted work)
    h.L.hh_force_general_path(1)
    res = {}
    for name, mode in (("torque", 0), ("pd", 1), ("jacobian", 2), ("osc", 3)):
        tot = []
        for phase in (0.0, 1.5, 3.0, 4.5):
            c = O.Cassie2d(m)
            part = 0
            for k in range(n_steps):
                s = c.op_state()
                a_j = squat_jacobian_action(s, k * 0.0005, phase); a_o = squat_osc_action(s, k * 0.0005, phase)
                q, v = c.data.state(); w = c.data.warmstart()
                act = {0: c.last_ctrl(), 1: QPOS_INIT_PY[[3, 4, 6, 8, 9, 11]], 2: a_j, 3: a_o}[mode]
                out = np.zeros(10, np.int64); out[8] = part
                h.L.hh_count_ops(mode, h.p(q), h.p(v), h.p(w), h.p(np.ascontiguousarray(act, np.float64)), len(act),
                                 out.ctypes.data_as(lp))
                part = int(out[9])
                fl = int(out[:4].sum())
                if mode == 3:
                    fl += QP_ASSEMBLY + QP_ITER * int(out[8])
                tot.append((fl, int(out[4]), int(out[6]), int(out[7]), int(out[8])))
                (c.step_osc(a_o) if mode == 3 else c.step_jacobian(a_j))
        t = np.array(tot)
        res[name] = t[:, 0].mean()
        print("%-9s flops/step mean %.0f (min %d max %d)  transcendental %.0f  rows %.1f  sweeps %.1f  qp iters %.2f" %
              (name, t[:, 0].mean(), t[:, 0].min(), t[:, 0].max(), t[:, 1].mean(), t[:, 2].mean(), t[:, 3].mean(), t[:, 4].mean()))
    return res


if __name__ == "__main__":
    main()
