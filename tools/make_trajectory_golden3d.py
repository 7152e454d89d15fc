"""tests/golden/traj3d_reference.npz: the REFERENCE's Cassie3dTraj / Cassie2dTraj (imported from
/root/reference/rllab/envs/cassie2d_trajectory.py) run on a small SYNTHETIC stepdata table that the test
rebuilds from the same seed -- state(t), action(t), sample() under random.seed, quat2eul.  Run in the build
container only (needs /root/reference)."""
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from traj_synth import synthetic_stepdata  # noqa: E402

sys.path.insert(0, os.path.join(REF, "rllab/envs"))
import cassie2d_trajectory as ref_mod  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "synth.bin")
    synthetic_stepdata().tofile(path)
    r3, r2 = ref_mod.Cassie3dTraj(path), ref_mod.Cassie2dTraj(path)
times = np.array([0.0, 0.013, 0.4, 0.77, 1.5, 9.99])
random.seed(11)
s3 = [r3.sample() for _ in range(6)]
random.seed(12)
s2 = [r2.sample() for _ in range(6)]
quats = np.random.default_rng(3).normal(size=(8, 4))
quats /= np.linalg.norm(quats, axis=1, keepdims=True)
golden = dict(
    times=times,
    st3_q=np.array([r3.state(t)[0] for t in times]), st3_v=np.array([r3.state(t)[1] for t in times]),
    ac3_mpos=np.array([r3.action(t)[0] for t in times]), ac3_mvel=np.array([r3.action(t)[1] for t in times]),
    ac3_tau=np.array([r3.action(t)[2] for t in times]),
    st2_q=np.array([r2.state(t)[0] for t in times]), st2_v=np.array([r2.state(t)[1] for t in times]),
    ac2_tau=np.array([r2.action(t)[2] for t in times]),
    q2=r2.qpos, v2=r2.qvel, tau2=r2.torque,
    s3_t=np.array([s[0] for s in s3]), s3_q=np.array([s[1] for s in s3]), s3_v=np.array([s[2] for s in s3]),
    s2_t=np.array([s[0] for s in s2]), s2_q=np.array([s[1] for s in s2]), s2_v=np.array([s[2] for s in s2]),
    quats=quats, eul=np.array([r2.quat2eul(*q) for q in quats]))
np.savez_compressed(os.path.join(ROOT, "tests/golden/traj3d_reference.npz"), **golden)
print("wrote tests/golden/traj3d_reference.npz")
