"""Builds cassierl_b200/model/stepdata_2d.npz from the reference's rllab/trajectory/stepdata.bin
with OUR loader, and tests/golden/traj2d_reference.npz with the REFERENCE's own loader
(rllab/envs/cassie2d_trajectory.py imported from /root/reference) so that the test can compare
the two.  Run in the build container only (needs /root/reference)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from cassierl_b200.trajectory import Cassie2dTraj as Ours  # noqa: E402

bin_path = os.path.join(REF, "rllab/trajectory/stepdata.bin")
ours = Ours(bin_path)
ours.save_npz(os.path.join(ROOT, "cassierl_b200/model/stepdata_2d.npz"))

sys.path.insert(0, os.path.join(REF, "rllab/envs"))
import cassie2d_trajectory as ref_mod  # noqa: E402

ref = ref_mod.Cassie2dTraj(bin_path)
times = np.array([0.0, 0.0005, 0.005, 0.01, 0.42, 0.84, 0.8405, 1.0, 5.0, 12.3456])
# the env's clock: self.time += 0.0005 ten times per policy step (cassie2d.py:122)
t = 0.0
idx_seq = []
for k in range(400):
    for _ in range(10):
        t += 0.0005
    tmax = ref.time[-1]
    idx_seq.append(int((t % tmax) / tmax * len(ref.time)))
golden = dict(
    shape=np.array(ref.qpos.shape), tmax=np.array(ref.time[-1]), rows=ref.qpos[::40].copy(), qvel_rows=ref.qvel[::40].copy(),
    torque_rows=ref.torque[::40].copy(), times=times, state_at=np.array([ref.state(x)[0] for x in times]),
    idx_seq=np.array(idx_seq),
    sha_qpos=np.frombuffer(hashlib.sha256(np.ascontiguousarray(ref.qpos).tobytes()).digest(), np.uint8),
    sha_qvel=np.frombuffer(hashlib.sha256(np.ascontiguousarray(ref.qvel).tobytes()).digest(), np.uint8))
os.makedirs(os.path.join(ROOT, "tests/golden"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "tests/golden/traj2d_reference.npz"), **golden)
print("ours==ref qpos:", np.array_equal(ours.qpos, ref.qpos), "qvel:", np.array_equal(ours.qvel, ref.qvel),
      "torque:", np.array_equal(ours.torque, ref.torque), ref.qpos.shape)
