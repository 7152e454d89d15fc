"""Reference-trajectory tooling against golden vectors produced by the REFERENCE's own
rllab/envs/cassie2d_trajectory.py on a synthetic stepdata table (tools/make_trajectory_golden3d.py)."""
import os
import random

import numpy as np
import pytest

from cassierl_b200.trajectory import Cassie2dTraj, Cassie3dTraj, quat2eul
from traj_synth import synthetic_stepdata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "traj3d_reference.npz"))


@pytest.fixture(scope="module")
def synth_path(tmp_path_factory):
    p = tmp_path_factory.mktemp("traj") / "synth.bin"
    synthetic_stepdata().tofile(p)
    return str(p)


def test_3d_table_lookup(golden, synth_path):
    t3 = Cassie3dTraj(synth_path)
    assert t3.qpos.shape == (57, 35) and t3.qvel.shape == (57, 32) and t3.torque.shape == (57, 10)
    for k, t in enumerate(golden["times"]):
        q, v = t3.state(float(t))
        mp, mv, tau = t3.action(float(t))
        assert np.array_equal(q, golden["st3_q"][k]) and np.array_equal(v, golden["st3_v"][k])
        assert np.array_equal(mp, golden["ac3_mpos"][k]) and np.array_equal(mv, golden["ac3_mvel"][k])
        assert np.array_equal(tau, golden["ac3_tau"][k])


def test_2d_projection_of_synthetic_table(golden, synth_path):
    t2 = Cassie2dTraj(synth_path)
    assert np.array_equal(t2.qpos, golden["q2"])   # arcsin / clip are the same numpy calls: bit-exact
    assert np.array_equal(t2.qvel, golden["v2"]) and np.array_equal(t2.torque, golden["tau2"])
    for k, t in enumerate(golden["times"]):
        assert np.array_equal(t2.state(float(t))[0], golden["st2_q"][k])
        assert np.array_equal(t2.state(float(t))[1], golden["st2_v"][k])
        assert np.array_equal(t2.action(float(t))[2], golden["ac2_tau"][k])


def test_random_phase_sample_follows_python_random(golden, synth_path):
    t3, t2 = Cassie3dTraj(synth_path), Cassie2dTraj(synth_path)
    random.seed(11)
    for k in range(6):
        tt, q, v = t3.sample()
        assert tt == golden["s3_t"][k] and np.array_equal(q, golden["s3_q"][k]) and np.array_equal(v, golden["s3_v"][k])
    random.seed(12)
    for k in range(6):
        tt, q, v = t2.sample()
        assert tt == golden["s2_t"][k] and np.array_equal(q, golden["s2_q"][k]) and np.array_equal(v, golden["s2_v"][k])


def test_quat2eul(golden):
    for q, e in zip(golden["quats"], golden["eul"]):
        assert np.allclose(np.array(quat2eul(*q)), e, rtol=0, atol=1e-15)
    # scalar edge: the reference clamps 2(wy - zx) into [-1, 1] before asin
    assert quat2eul(np.sqrt(0.5), 0.0, np.sqrt(0.5) + 1e-9, 0.0)[1] == pytest.approx(np.pi / 2)
