"""Reference-trajectory loader for the imitation reward.

Restates rllab/envs/cassie2d_trajectory.py (CassieRL/cassierl): `stepdata.bin` wire format
(:5-14: rows of 1 time + 35 qpos + 32 qvel + 10 torque + 10 mpos + 10 mvel float64), the
3-D -> 2-D projection (:31-134: base/rod quaternions -> ZYX Euler Y angle, drop the abduction,
yaw, spring and rod-hinge columns) and the time -> row lookup (:16-19), vectorised.
"""
import os
import random

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PACKAGED_TABLE = os.path.join(_HERE, "model", "stepdata_2d.npz")

ROW = 1 + 35 + 32 + 10 + 10 + 10
# 3-D qpos columns that survive (after the quaternion's x slot has been overwritten by Euler Y):
# base x, z, pitch | per leg: hip, knee, ankle, toe, rod pitch
QPOS_KEEP = [0, 2, 4, 9, 10, 12, 13, 18, 23, 24, 26, 27, 32]
QVEL_KEEP = [0, 2, 4, 8, 9, 11, 12, 17, 21, 22, 24, 25, 30]
TORQUE_KEEP = [2, 3, 4, 7, 8, 9]


def _euler_y(w, x, y, z):
    """middle angle of the reference's quat2eul (:136-151): asin(clip(2(wy - zx)))"""
    return np.arcsin(np.clip(2.0 * (w * y - z * x), -1.0, 1.0))


def quat2eul(w, x, y, z):
    """ZYX Euler angles (Z, Y, X) of a quaternion, the reference's Cassie2dTraj.quat2eul (:136-151)."""
    X = np.arctan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    Y = _euler_y(w, x, y, z)
    Z = np.arctan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return Z, Y, X


class Cassie3dTraj:
    """The raw 3-D table of `stepdata.bin` (reference class of the same name, :5-28): .time .qpos(35) .qvel(32)
    .torque(10) .mpos(10) .mvel(10), state(t), action(t), sample()."""

    def __init__(self, filepath):
        data = np.fromfile(filepath, dtype=np.float64).reshape((-1, ROW))
        self.time = data[:, 0].copy()
        self.qpos = data[:, 1:36].copy()
        self.qvel = data[:, 36:68].copy()
        self.torque = data[:, 68:78].copy()
        self.mpos = data[:, 78:88].copy()
        self.mvel = data[:, 88:98].copy()

    def index(self, t):
        tmax = self.time[-1]
        return int((t % tmax) / tmax * len(self.time))

    def state(self, t):
        i = self.index(t)
        return (self.qpos[i], self.qvel[i])

    def action(self, t):
        i = self.index(t)
        return (self.mpos[i], self.mvel[i], self.torque[i])

    def sample(self):
        """random-phase draw (:26-28): uses the `random` module like the reference, so `random.seed` reproduces it"""
        i = random.randrange(len(self.time))
        return (self.time[i], self.qpos[i], self.qvel[i])


class Cassie2dTraj:
    """Same interface as the reference class: .time .qpos .qvel .torque, state(t), action(t), sample()."""

    def __init__(self, filepath=None):
        if filepath is None or filepath.endswith(".npz"):
            z = np.load(filepath or PACKAGED_TABLE)
            self.time, self.qpos, self.qvel, self.torque = z["time"], z["qpos"], z["qvel"], z["torque"]
            self.mpos, self.mvel = z["mpos"], z["mvel"]
            return
        data = np.fromfile(filepath, dtype=np.float64).reshape((-1, ROW))
        self.time = data[:, 0].copy()
        q3 = data[:, 1:36].copy()
        for c in (3, 17, 31):  # base, left rod, right rod quaternions (w x y z)
            q3[:, c + 1] = _euler_y(q3[:, c], q3[:, c + 1], q3[:, c + 2], q3[:, c + 3])
        self.qpos = np.ascontiguousarray(q3[:, QPOS_KEEP])
        self.qvel = np.ascontiguousarray(data[:, 36:68][:, QVEL_KEEP])
        self.torque = np.ascontiguousarray(data[:, 68:78][:, TORQUE_KEEP])
        self.mpos = data[:, 78:88].copy()
        self.mvel = data[:, 88:98].copy()

    def index(self, t):
        tmax = self.time[-1]
        return int((t % tmax) / tmax * len(self.time))

    def state(self, t):
        i = self.index(t)
        return (self.qpos[i], self.qvel[i])

    def action(self, t):
        i = self.index(t)
        return (self.mpos[i], self.mvel[i], self.torque[i])

    def sample(self):
        i = random.randrange(len(self.time))
        return (self.time[i], self.qpos[i], self.qvel[i])

    def quat2eul(self, w, x, y, z):
        return quat2eul(w, x, y, z)

    def save_npz(self, path):
        np.savez_compressed(path, time=self.time, qpos=self.qpos, qvel=self.qvel, torque=self.torque,
                            mpos=self.mpos, mvel=self.mvel)
