"""The quad engine's constraint tiers (cassierl_b200/csrc/quad_engine.cuh compiled for the CPU by
tests/host_harness/quad_harness.cpp: the four lanes of an env run as four threads that meet at a barrier for every
shuffle) against the fp64 oracle, without a GPU.

Tier 0 = 12 rows (standing / squatting), tier 1 = 16 rows (two joint limits per leg), tier 2 = 20 rows (every limited
joint of a leg at its stop: robots in flight under random OSC accelerations, rllab/envs/cassie_stand2d.py:50 action
space), anything else = serial fallback.  All tiers restate the same mj_step [EXT] (Cassie2d.cpp:92), so every tier
must agree with the oracle to the fp64 bar (1e-9 relative per step) and with each other."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

from conftest import QPOS_INIT_PY, ROOT, TORQUE_HIGH, _stale, rel_err


class QuadHarness:
    def __init__(self, path, xml):
        self.L = ct.CDLL(path)
        self.L.qh_error.restype = ct.c_char_p
        assert self.L.qh_load(xml.encode()) == 0, self.L.qh_error()

    def force(self, tier1=False, tier2=False, general=False):
        self.L.qh_force_tier1(int(tier1)); self.L.qh_force_tier2(int(tier2)); self.L.qh_force_general_path(int(general))

    def ctrl_steps(self, mode, q, qd, warm, act, f32=False, qp_set=0):
        """n controller steps (mode 0 torque, 1 pd, 2 jacobian, 3 osc) through quad_controller_step; same outputs as
        conftest.Harness.ctrl_steps plus the QP partition carried from step to step"""
        dp = ct.POINTER(ct.c_double)
        act = np.ascontiguousarray(act, np.float64)
        n, adim = act.shape
        u = np.zeros((n, 6)); op = np.zeros((n, 18)); traj = np.zeros((n, 26)); mk = np.zeros(n, np.uint32)
        qp = np.zeros((n, 2), np.int32); qs = ct.c_uint(qp_set)
        fn = self.L.qh_ctrl_steps_f32 if f32 else self.L.qh_ctrl_steps_f64
        fn(int(mode), n, q.ctypes.data_as(dp), qd.ctypes.data_as(dp), warm.ctypes.data_as(dp), act.ctypes.data_as(dp), adim,
           u.ctypes.data_as(dp), op.ctypes.data_as(dp), traj.ctypes.data_as(dp), mk.ctypes.data_as(ct.POINTER(ct.c_uint)),
           qp.ctypes.data_as(ct.POINTER(ct.c_int)), ct.byref(qs))
        return dict(u=u, op=op, traj=traj, mask=mk, qp=qp, qp_set=qs.value)

    def steps(self, q, qd, warm, u, f32=False):
        dp = ct.POINTER(ct.c_double); ip = ct.POINTER(ct.c_int); up = ct.POINTER(ct.c_uint)
        u = np.ascontiguousarray(u, np.float64).reshape(-1, 6)
        n = u.shape[0]
        traj = np.zeros((n, 26)); nr = np.zeros(n, np.int32); sw = np.zeros(n, np.int32); mk = np.zeros(n, np.uint32)
        fn = self.L.qh_steps_f32 if f32 else self.L.qh_steps_f64
        for k in range(n):
            fn(1, q.ctypes.data_as(dp), qd.ctypes.data_as(dp), warm.ctypes.data_as(dp), u[k].ctypes.data_as(dp),
               nr[k:].ctypes.data_as(ip), sw[k:].ctypes.data_as(ip), mk[k:].ctypes.data_as(up))
            traj[k, :13] = q; traj[k, 13:] = qd
        return traj, nr, sw, mk


@pytest.fixture(scope="module")
def qharness(oracle):
    src = os.path.join(ROOT, "tests", "host_harness", "quad_harness.cpp")
    csrc = os.path.join(ROOT, "cassierl_b200", "csrc")
    out = os.path.join(ROOT, "tests", "_build", "libquad_harness.so")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if _stale(out, deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", out, src,
                               os.path.join(csrc, "mjcf_flatten.cpp")])
    h = QuadHarness(out, oracle.default_model_path())
    yield h
    h.force()


def oracle_steps(oracle, omodel, q0, qd0, u):
    d = oracle.Data(omodel)
    d.set_state(q0, qd0)
    traj = []; masks = []
    for k in range(len(u)):
        d.step(u[k])
        q, v = d.state()
        traj.append(np.concatenate([q, v])); masks.append(d.contact_mask())
    return np.array(traj), np.array(masks, np.uint64)


def limited_joints(omodel):
    """(dof, lo, hi) of the limited hinges, from the planar file: hip, knee, tarsus ('ankle'), toe of each leg"""
    from oracle import mjcf_reader
    from oracle.oracle import default_model_path
    js = mjcf_reader.read_mjcf(default_model_path())["joints"]
    return [(j, float(js[j]["range"][0]), float(js[j]["range"][1])) for j in range(13) if js[j]["limited"]]


def flight_state(omodel, seed, nl_left, nl_right, z=1.6):
    """robot in the air, nl_left / nl_right limited joints of the two legs pushed past a stop"""
    rng = np.random.default_rng(seed)
    q = QPOS_INIT_PY.copy(); q[1] = z
    qd = 0.5 * rng.standard_normal(13)
    lim = limited_joints(omodel)
    left = [t for t in lim if t[0] < 8]; right = [t for t in lim if t[0] >= 8]
    for side, n in ((left, nl_left), (right, nl_right)):
        for j, lo, hi in [side[i] for i in rng.permutation(len(side))[:n]]:
            q[j] = hi + rng.uniform(0.005, 0.03) if rng.random() < 0.5 else lo - rng.uniform(0.005, 0.03)
    return q, qd


def test_tiers_agree_on_the_common_regime(qharness, oracle, omodel):
    """standing robot under random torques: tier 0, forced tier 1 and forced tier 2 give the oracle's trajectory"""
    rng = np.random.default_rng(5)
    n = 60
    u = np.repeat(rng.uniform(-1, 1, (n // 10, 6)) * TORQUE_HIGH, 10, axis=0)
    ref, masks = oracle_steps(oracle, omodel, QPOS_INIT_PY, np.zeros(13), u)
    out = {}
    for name, kw in (("t0", {}), ("t1", dict(tier1=True)), ("t2", dict(tier2=True))):
        qharness.force(**kw)
        q = QPOS_INIT_PY.copy(); qd = np.zeros(13); w = np.zeros(13)
        traj, nr, sw, mk = qharness.steps(q, qd, w, u)
        assert rel_err(traj, ref) < 1e-9, name
        assert np.array_equal(mk.astype(np.uint64), masks), name
        out[name] = traj
    qharness.force()
    # inert rows contribute exact zeros: the tiers agree far below the oracle bar
    assert np.abs(out["t1"] - out["t0"]).max() < 1e-12 and np.abs(out["t2"] - out["t0"]).max() < 1e-12


@pytest.mark.parametrize("nl", [(3, 0), (4, 0), (3, 3), (4, 4), (4, 1), (2, 3)])
def test_tier2_joint_limits_in_flight(qharness, oracle, omodel, nl):
    """3 or 4 joint limits on a leg (previously the serial fallback): the cooperative 20-row tier == oracle, 1e-9"""
    qharness.force()
    for seed in range(4):
        q0, qd0 = flight_state(omodel, 100 * nl[0] + 10 * nl[1] + seed, *nl)
        u = np.zeros((5, 6))
        ref, masks = oracle_steps(oracle, omodel, q0, qd0, u)
        q = q0.copy(); qd = qd0.copy(); w = np.zeros(13)
        traj, nr, sw, mk = qharness.steps(q, qd, w, u)
        assert nr[0] == 4 + nl[0] + nl[1]
        assert rel_err(traj, ref) < 1e-9
        # ... and the serial fallback (thread engine) on the same state
        qharness.force(general=True)
        q = q0.copy(); qd = qd0.copy(); w = np.zeros(13)
        traj_g, nr_g, _, _ = qharness.steps(q, qd, w, u)
        qharness.force()
        assert np.array_equal(nr, nr_g)
        assert rel_err(traj, traj_g) < 1e-9


def test_tier2_limits_with_contacts(qharness, oracle, omodel):
    """landing with the legs at their stops: limits and toe contacts in the same solve"""
    qharness.force()
    hit = 0
    for seed in range(6):
        q0, qd0 = flight_state(omodel, 900 + seed, 3, 4, z=QPOS_INIT_PY[1] + 0.002)
        qd0[1] = -0.5
        u = np.zeros((20, 6))
        ref, masks = oracle_steps(oracle, omodel, q0, qd0, u)
        q = q0.copy(); qd = qd0.copy(); w = np.zeros(13)
        traj, nr, sw, mk = qharness.steps(q, qd, w, u)
        assert rel_err(traj, ref) < 1e-9
        assert np.array_equal(mk.astype(np.uint64), masks)
        hit += int(((nr > 11) & (mk != 0)).any())
    assert hit > 0


def test_tier2_fp32_single_step(qharness, oracle, omodel):
    """fp32 build of the 20-row tier, one step from the oracle's state: 1e-5 relative (BASELINE.json tolerance)"""
    qharness.force()
    worst = 0.0
    for seed in range(6):
        q0, qd0 = flight_state(omodel, 40 + seed, 4, 3)
        u = np.zeros((1, 6))
        ref, _ = oracle_steps(oracle, omodel, q0, qd0, u)
        q = q0.copy(); qd = qd0.copy(); w = np.zeros(13)
        traj, nr, _, _ = qharness.steps(q, qd, w, u, f32=True)
        assert nr[0] == 11
        worst = max(worst, rel_err(traj, ref))
    assert worst < 1e-5, worst


# ----------------------------------------------------------------------------- the cooperative controllers
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_quad_controllers_match_oracle_facade(qharness, oracle, omodel, mode):
    """StepPd / StepJacobian / StepOsc (Cassie2d.cpp:96-209, OSC_RBDL.cpp:114-291) as the quad engine runs them -- the task
    loop, the 14-variable box QP and the pseudo-inverses split over four lanes -- against the oracle facade: free-running
    fp64 trajectory, torques and lagged op-space state."""
    from conftest import QPOS_INIT_CTOR, squat_jacobian_action, squat_osc_action
    from test_engine_host import run_facade
    qharness.force()
    n = 60
    if mode == 1:
        rng = np.random.default_rng(7)
        tg = QPOS_INIT_CTOR[[3, 4, 6, 8, 9, 11]] + rng.uniform(-0.1, 0.1, (n // 10, 6))
        fn = lambda k, c: tg[k // 10]
    elif mode == 2:
        fn = lambda k, c: squat_jacobian_action(c.op_state(), k * 0.0005)
    else:
        fn = lambda k, c: squat_osc_action(c.op_state(), k * 0.0005)
    ref, us, ops, acts, starts = run_facade(oracle, omodel, mode, fn, n)
    q = QPOS_INIT_CTOR.copy(); qd = np.zeros(13); w = np.zeros(13)
    out = qharness.ctrl_steps(mode, q, qd, w, acts)
    assert rel_err(out["traj"], ref) < (1e-9 if mode != 3 else 1e-8)
    assert rel_err(out["u"], us) < (1e-8 if mode != 3 else 1e-6)
    assert rel_err(out["op"], ops) < 1e-8
    if mode == 3:
        assert out["qp"][:, 1].max() == 0          # every QP solved to optimality
        assert np.median(out["qp"][:, 0]) == 1      # warm partition: one KKT solve per step


def test_quad_random_torque_trajectory_crosses_tiers(qharness, oracle, omodel):
    """config 2's stream on the quad engine: 2000 free-running steps of random torques (the robot is thrown around and
    ends up on the floor: joint-limit rows and body contacts come and go, so the steps alternate between the cooperative
    tiers and the serial fallback) == oracle, contact events identical."""
    from test_engine_host import oracle_torque_traj, random_torques
    qharness.force()
    n = 2000
    u = random_torques(n, 3)
    ref, masks, _ = oracle_torque_traj(oracle, omodel, QPOS_INIT_PY, u)
    q = QPOS_INIT_PY.copy(); qd = np.zeros(13); w = np.zeros(13)
    traj, nr, sw, mk = qharness.steps(q, qd, w, u)
    assert rel_err(traj, ref) < 1e-9
    assert np.array_equal(mk.astype(np.uint64), masks)
    assert nr.max() > 12      # the stream left the common tier


@pytest.mark.parametrize("mode", [2, 3])
def test_quad_controllers_fp32_single_step(qharness, oracle, omodel, mode):
    """fp32 physics + double-precision controller (the product's fp32 build) on the quad engine, teacher-forced from the
    oracle's states along the squatting stream: 1e-5 relative on the state (BASELINE.json tolerance)."""
    from conftest import squat_jacobian_action, squat_osc_action
    from test_engine_host import run_facade
    qharness.force()
    n = 200
    act = squat_jacobian_action if mode == 2 else squat_osc_action
    ref, us, ops, acts, starts = run_facade(oracle, omodel, mode, lambda k, c: act(c.op_state(), k * 0.0005), n)
    # the QP partition is carried like the kernel carries it: from the fp64 run of the same stream
    worst = 0.0
    qp_set = 0
    for k in range(0, n):
        q, v, ws = [x.copy() for x in starts[k]]
        o = qharness.ctrl_steps(mode, q, v, ws, acts[k:k + 1], f32=True, qp_set=qp_set)
        qp_set = o["qp_set"]
        if k % 4 == 0:
            worst = max(worst, rel_err(o["traj"][0], ref[k]))
    assert worst < 1e-5, worst
