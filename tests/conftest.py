import ctypes as ct
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


@pytest.fixture(scope="session")
def oracle():
    """The fp64 CPU oracle (oracle/), built with gcc on first use."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def omodel(oracle):
    return oracle.Model()


class Harness:
    """ctypes view of tests/host_harness (the device headers compiled for the CPU)."""

    def __init__(self, path, xml):
        self.L = ct.CDLL(path)
        self.L.hh_error.restype = ct.c_char_p
        self.L.hh_total_mass.restype = ct.c_double
        assert self.L.hh_load(xml.encode()) == 0, self.L.hh_error()

    @staticmethod
    def p(a):
        return a.ctypes.data_as(ct.POINTER(ct.c_double))

    def dynamics(self, ctrl, q, qd):
        M = np.zeros((13, 13)); b = np.zeros(13)
        self.L.hh_dynamics(int(ctrl), self.p(np.ascontiguousarray(q, np.float64)), self.p(np.ascontiguousarray(qd, np.float64)),
                           self.p(M), self.p(b))
        return M, b

    def steps(self, q, qd, warm, u, f32=False):
        """n torque steps; returns traj [n,26], nrows, sweeps, mask (q, qd, warm updated in place)"""
        u = np.ascontiguousarray(u, np.float64).reshape(-1, 6)
        n = u.shape[0]
        traj = np.zeros((n, 26)); nr = np.zeros(n, np.int32); sw = np.zeros(n, np.int32); mk = np.zeros(n, np.uint32)
        fn = self.L.hh_steps_f32 if f32 else self.L.hh_steps_f64
        ip = ct.POINTER(ct.c_int); up = ct.POINTER(ct.c_uint)
        for k in range(n):
            fn(1, self.p(q), self.p(qd), self.p(warm), self.p(u[k]), nr[k:].ctypes.data_as(ip), sw[k:].ctypes.data_as(ip),
               mk[k:].ctypes.data_as(up))
            traj[k, :13] = q; traj[k, 13:] = qd
        return traj, nr, sw, mk

    def ctrl_steps(self, mode, q, qd, warm, act, f32=False):
        act = np.ascontiguousarray(act, np.float64)
        n, adim = act.shape
        u = np.zeros((n, 6)); op = np.zeros((n, 18)); traj = np.zeros((n, 26)); mk = np.zeros(n, np.uint32)
        qp = np.zeros((n, 2), np.int32)
        fn = self.L.hh_ctrl_steps_f32 if f32 else self.L.hh_ctrl_steps_f64
        fn(int(mode), n, self.p(q), self.p(qd), self.p(warm), self.p(act), adim, self.p(u), self.p(op), self.p(traj),
           mk.ctypes.data_as(ct.POINTER(ct.c_uint)), qp.ctypes.data_as(ct.POINTER(ct.c_int)))
        return dict(u=u, op=op, traj=traj, mask=mk, qp=qp)

    def squat(self, mode, n, phase, q, qd, warm, f32=False):
        traj = np.zeros((n, 26)); u = np.zeros((n, 6)); qp = np.zeros((n, 2), np.int32)
        fn = self.L.hh_squat_f32 if f32 else self.L.hh_squat_f64
        fn(int(mode), n, ct.c_double(phase), self.p(q), self.p(qd), self.p(warm), self.p(traj), self.p(u),
           qp.ctypes.data_as(ct.POINTER(ct.c_int)))
        self.last_qp = qp
        return traj, u

    def ctrl_dynamics(self, q, qd):
        bias = np.zeros(13); Jeq = np.zeros((4, 13)); gamma = np.zeros(13); Nc = np.zeros((13, 13))
        self.L.hh_ctrl_dynamics(self.p(np.ascontiguousarray(q, np.float64)), self.p(np.ascontiguousarray(qd, np.float64)),
                                self.p(bias), self.p(Jeq), self.p(gamma), self.p(Nc))
        return bias, Jeq, gamma, Nc


@pytest.fixture(scope="session")
def harness(oracle):
    src = os.path.join(ROOT, "tests", "host_harness", "harness.cpp")
    csrc = os.path.join(ROOT, "cassierl_b200", "csrc")
    out = os.path.join(ROOT, "tests", "_build", "libhost_harness.so")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if _stale(out, deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src,
                               os.path.join(csrc, "mjcf_flatten.cpp")])
    return Harness(out, oracle.default_model_path())


class TreeHarness:
    """ctypes view of tests/host_harness/tree_harness.cpp: the 3-D tree engine as a one-lane tile on the CPU."""

    def __init__(self, path, xml):
        self.L = ct.CDLL(path)
        self.L.th_error.restype = ct.c_char_p
        assert self.L.th_load(xml.encode()) == 0, self.L.th_error()
        sz = (ct.c_int * 10)()
        self.L.th_sizes(sz)
        (self.nl, self.nv, self.nq, self.nu, self.ng, self.npair, self.neq, self.nlevels, self.scratch32, self.scratch64) = list(sz)

    p = staticmethod(Harness.p)

    @staticmethod
    def ip(a):
        return a.ctypes.data_as(ct.POINTER(ct.c_int))

    def consts(self):
        diw = np.zeros(self.nv); mi = ct.c_double(); piw = np.zeros(self.npair); eiw = np.zeros(self.neq)
        q0 = np.zeros(self.nq); lm = np.zeros(self.nl)
        self.L.th_consts(self.p(diw), ct.byref(mi), self.p(piw), self.p(eiw), self.p(q0), self.p(lm))
        return dict(dof_invweight0=diw, meaninertia=mi.value, pair_invweight=piw, eq_invweight=eiw, qpos0=q0, link_mass=lm)

    def dynamics(self, q, qd):
        M = np.zeros((self.nv, self.nv)); b = np.zeros(self.nv)
        self.L.th_dynamics(self.p(np.ascontiguousarray(q, np.float64)), self.p(np.ascontiguousarray(qd, np.float64)), self.p(M), self.p(b))
        return M, b

    def rows(self, q, qd):
        J = np.zeros((48, self.nv)); pos = np.zeros(48); R = np.zeros(48); aref = np.zeros(48)
        tp = np.zeros(48, np.int32); idd = np.zeros(48, np.int32)
        n = self.L.th_rows(self.p(np.ascontiguousarray(q, np.float64)), self.p(np.ascontiguousarray(qd, np.float64)), self.p(J),
                           self.p(pos), self.p(R), self.p(aref), self.ip(tp), self.ip(idd))
        return dict(J=J[:n], pos=pos[:n], R=R[:n], aref=aref[:n], type=tp[:n], id=idd[:n])

    def step(self, q, qd, warm, u, f32=False, n=1):
        """n steps in place on (q, qd, warm); returns stats (rows, contacts, sweeps, dropped)"""
        st = np.zeros(4, np.int32)
        fn = self.L.th_steps_f32 if f32 else self.L.th_steps_f64
        fn(int(n), self.p(q), self.p(qd), self.p(warm), self.p(np.ascontiguousarray(u, np.float64)), self.ip(st))
        return st

    def step_fast(self, q, qd, warm, u):
        """the kernels' first pass: fast capacity, abort on overflow.  Returns (taken, stats)"""
        st = np.zeros(4, np.int32)
        ok = self.L.th_step_fast_f64(self.p(q), self.p(qd), self.p(warm), self.p(np.ascontiguousarray(u, np.float64)), self.ip(st))
        return bool(ok), st

    def count_ops(self, q, qd, warm, u):
        out = np.zeros(6, np.int64)
        self.L.th_count_ops(self.p(np.ascontiguousarray(q, np.float64)), self.p(np.ascontiguousarray(qd, np.float64)),
                            self.p(np.ascontiguousarray(warm, np.float64)), self.p(np.ascontiguousarray(u, np.float64)),
                            out.ctypes.data_as(ct.POINTER(ct.c_long)))
        return out


@pytest.fixture(scope="session")
def omodel3d(oracle):
    return oracle.Model(oracle.model3d_path())


@pytest.fixture(scope="session")
def tree_harness(oracle):
    src = os.path.join(ROOT, "tests", "host_harness", "tree_harness.cpp")
    csrc = os.path.join(ROOT, "cassierl_b200", "csrc")
    out = os.path.join(ROOT, "tests", "_build", "libtree_harness.so")
    deps = [src] + [os.path.join(csrc, f) for f in ("tree_engine.cuh", "tree_model.h", "mjcf_flatten.cpp", "mjcf_flatten.h")]
    if _stale(out, deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src, os.path.join(csrc, "mjcf_flatten.cpp")])
    return TreeHarness(out, oracle.model3d_path())


LEG_POSE_3D = [0.0, 0.0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407]
TORQUE_HIGH_3D = np.array([4.5, 4.5, 12.2, 12.2, 0.9] * 2)


def lying3d(z=0.10, axis="x", angle=1.5708, straight=False):
    """a robot pressed into the floor on its side / back: 9-10 contacts, 33-36 constraint rows (more than the fast capacity)"""
    q = pose3d(z)
    if straight:
        q[7:14] = [0, 0, 0.0, -0.7, 1.0, -1.5, 0]; q[14:21] = q[7:14]
    v = [np.cos(angle / 2), 0.0, 0.0, 0.0]
    v[{"x": 1, "y": 2}[axis]] = np.sin(angle / 2)
    q[3:7] = v
    return q


def pose3d(z=0.945):
    """standing pose of Cassie2d.cpp:56-58 on the 3-D joints (abduction = yaw = 0); the toes touch the floor at z = 0.935"""
    q = np.zeros(21)
    q[2] = z; q[3] = 1.0
    q[7:14] = LEG_POSE_3D; q[14:21] = LEG_POSE_3D
    return q


QPOS_INIT_PY = np.array([0.0, 0.939, 0.0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
                         0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407])
QPOS_INIT_CTOR = np.array([0.0, 0.939, 0.0, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407,
                           0.68111815, -1.40730353, 1.62972043, -1.77611107, -0.61968402])
TORQUE_HIGH = np.array([12.0, 12.0, 0.9, 12.0, 12.0, 0.9])


def rel_err(a, b):
    """max |a-b| / max(1, |b|): RELATIVE for |value| >= 1 and ABSOLUTE below 1 (most joint angles and
    velocities of this robot are below 1, so for them the 1e-5 / 1e-9 bars of the parity tests are absolute
    bounds in rad, m, rad/s, m/s).  DESIGN.md section 7 states this next to every tolerance."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def norm_rel_err(a, b):
    """||a-b||_inf / ||b||_inf of one vector (qpos or qvel): the true norm-wise relative error."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b))) / d if d > 0 else float(np.max(np.abs(a - b)))


def squat_jacobian_action(s, t, phase=0.0):
    """standing_controller_jacobian (rllab/envs/cassie2d.py:297-331) + squatting.py:8-16 targets"""
    w = 0.5 * 3.1415
    zt = 0.7 + 0.25 * np.sin(w * t + phase); zdt = 0.25 * np.cos(w * t + phase)
    xt = (s[6] + s[12]) / 2.0
    fx = 200.0 * (xt - s[0]) + 50.0 * (0.0 - s[3])
    fz = 0.5 * 9.806 * 31.0 + 200.0 * (zt - s[1]) + 50.0 * (zdt - s[4])
    my = 100.0 * (0.0 - s[2]) + 10.0 * (0.0 - s[5])
    fz = max(fz, 0.0)
    return np.array([fx, fz, my, fx, fz, my])


def squat_osc_action(s, t, phase=0.0):
    """standing_controller_osc (rllab/envs/cassie2d.py:263-295)"""
    w = 0.5 * 3.1415
    zt = 0.7 + 0.25 * np.sin(w * t + phase); zdt = 0.25 * np.cos(w * t + phase)
    xt = (s[6] + s[12]) / 2.0
    return np.array([100.0 * (xt - s[0]) + 20.0 * (0.0 - s[3]), 100.0 * (zt - s[1]) + 20.0 * (zdt - s[4]),
                     0.0, 100.0 * (-5e-3 - s[7]), 0.0, 100.0 * (-5e-3 - s[13]),
                     20.0 * (0.0 - s[2]) + 10.0 * (0.0 - s[5])])
