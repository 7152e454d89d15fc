"""Synthetic `stepdata.bin` table (wire format of rllab/envs/cassie2d_trajectory.py:5-14) shared by
tools/make_trajectory_golden3d.py and tests/test_trajectory.py."""
import numpy as np


def synthetic_stepdata(rows=57, seed=5):
    rng = np.random.default_rng(seed)
    data = rng.normal(size=(rows, 1 + 35 + 32 + 10 + 10 + 10))
    data[:, 0] = np.cumsum(rng.uniform(0.01, 0.03, size=rows))
    for c in (3, 17, 31):  # base and connecting-rod quaternions (w x y z) inside qpos
        q = data[:, 1 + c:1 + c + 4]
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(data)
