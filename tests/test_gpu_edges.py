"""Edge cases of the batch ABI on the GPU: ragged / tiny batches, zero substeps, masked resets, the
cold constraint path (joint limits, body-on-floor contacts) in fp32, error reporting, imitation-task
variants.  Same oracle and tolerances as test_gpu_parity.py."""
import ctypes as ct

import numpy as np
import pytest

from conftest import QPOS_INIT_CTOR, QPOS_INIT_PY, TORQUE_HIGH, rel_err
from test_gpu_parity import q_from_s26, s26

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from cassierl_b200 import envs
    return envs


@pytest.fixture(scope="module")
def LIB():
    from cassierl_b200 import lib
    return lib


@pytest.mark.parametrize("n", [1, 31, 33, 100])
def test_ragged_batches(E, LIB, oracle, omodel, n):
    """Batch sizes that do not fill a warp / a grid: every env still matches the oracle."""
    rng = np.random.default_rng(n)
    U = rng.uniform(-1, 1, (n, 3, 6)) * TORQUE_HIGH
    b = E.Cassie2dBatch(n, precision=64)
    for k in range(3):
        b.step_torque(torch.tensor(U[:, k]), 10)
    got = b.get_general_state().cpu().numpy()
    for e in range(0, n, max(1, n // 7)):
        d = oracle.Data(omodel); d.set_state(QPOS_INIT_CTOR, np.zeros(13))
        for k in range(3):
            for _ in range(10):
                d.step(U[e, k])
        q, v = d.state()
        assert rel_err(got[e], s26(oracle, q, v)) < 1e-9, e
    b.close()


def test_zero_substeps_and_masked_reset(E, LIB, oracle):
    n = 64
    b = E.Cassie2dBatch(n, precision=32)
    b.step_torque(torch.zeros(n, 6), 7)
    s0 = b.get_general_state().clone(); o0 = b.get_operational_space_state().clone(); w0 = b.get_warm_start().clone()
    b.step_torque(torch.ones(n, 6), 0)                       # n_substeps = 0: nothing moves
    assert torch.equal(b.get_general_state(), s0) and torch.equal(b.get_operational_space_state(), o0)
    mask = torch.zeros(n, dtype=torch.uint8); mask[::3] = 1
    st = s26(oracle, QPOS_INIT_PY, np.zeros(13)); st[1] = 1.25
    b.reset(st, mask=mask)                                   # Reset (Cassie2d.cpp:78-82) on every third env
    s1 = b.get_general_state()
    assert torch.equal(s1[1::3], s0[1::3]) and torch.equal(s1[2::3], s0[2::3])
    assert torch.allclose(s1[::3, 1], torch.tensor(1.25, device=s1.device))
    # Reset touches neither the solver warm start nor the lagged op-space state
    assert torch.equal(b.get_warm_start(), w0)
    o1 = b.get_operational_space_state()
    assert torch.equal(o1[:, [0, 1, 3, 4]], o0[:, [0, 1, 3, 4]])
    b.close()


def test_cold_path_fp32_single_step(E, LIB, oracle, omodel):
    """States with violated joint limits and tarsus / shin floor contacts (robots thrown around by
    random torques): the joint-limit / any-geom tiers of the constraint solver in fp32, one step from
    identical fp32-representable states on both sides (test_gpu_fp32_bar.py protocol), 1e-5, masks bit-exact."""
    n = 24
    rng = np.random.default_rng(77)
    f32 = lambda x: np.asarray(x, np.float64).astype(np.float32).astype(np.float64)
    pres = []; refs = []; acts = []; rmask = []; n_rare = 0
    for e in range(n):
        d = oracle.Data(omodel); d.set_state(QPOS_INIT_CTOR, np.zeros(13))
        U = rng.uniform(-1, 1, (80, 6)) * TORQUE_HIGH
        for k in range(500 + 10 * e):
            d.step(U[k // 10])
        (q, v), w = d.state(), d.warmstart()
        q, v, w = f32(q), f32(v), f32(w)
        a = f32(U[(500 + 10 * e) // 10])
        d.set_state(q, v); d.set_warmstart(w)
        d.step(a)
        ef = d.efc(); geoms = d.contacts()["geom"]           # geoms 5 / 9 are the toe capsules
        n_rare += bool((ef["type"] == 1).any() or any(g not in (5, 9) for g in geoms))
        q1, v1 = d.state()
        pres.append((s26(oracle, q, v), w)); refs.append(np.concatenate([q1, v1])); acts.append(a)
        rmask.append(d.contact_mask())
    assert n_rare >= 8                                       # joint-limit rows and / or tarsus, shin, thigh contacts
    b = E.Cassie2dBatch(n, precision=32)
    b.reset(torch.tensor(np.array([p[0] for p in pres]), dtype=torch.float32, device=b.device))
    b.set_warm_start(torch.tensor(np.array([p[1] for p in pres])))
    mask = torch.zeros(n, dtype=torch.int32, device=b.device)
    b.step_torque(torch.tensor(np.array(acts)), 1, contact_mask=mask)
    q, v = q_from_s26(b.get_general_state().cpu().numpy().astype(np.float64))
    err = np.abs(np.concatenate([q, v], axis=1) - np.array(refs)) / np.maximum(1, np.abs(np.array(refs)))
    print("cold path fp32: worst %.2e, median of per-env worst %.2e" % (err.max(), np.median(err.max(axis=1))))
    assert err.max() < 1e-5, err.max()
    assert np.array_equal(mask.cpu().numpy().astype(np.uint64), np.array(rmask, np.uint64))
    st = b.stats().cpu().numpy()
    assert (st[:, 0] >= 4).all() and (st[:, 1] <= 50).all()
    b.close()


def test_error_reporting(E, LIB):
    L = LIB.load()
    b = E.Cassie2dBatch(8)
    a = torch.zeros(8, 7, device=b.device)
    assert L.Cassie2dBatchStep(b.h, 9, a.data_ptr(), 1, None, None) < 0 and b"bad mode" in L.CassieGetLastError()
    assert L.Cassie2dBatchStep(b.h, 0, None, 1, None, None) < 0 and b"null" in L.CassieGetLastError()
    assert L.Cassie2dBatchStep(b.h, 0, a.data_ptr(), -1, None, None) < 0
    obs = torch.zeros(8, 26, device=b.device); r = torch.zeros(8, device=b.device); d = torch.zeros(8, dtype=torch.uint8, device=b.device)
    assert L.Cassie2dBatchEnvStep(b.h, 1, 1, a.data_ptr(), 10, 0, obs.data_ptr(), r.data_ptr(), d.data_ptr(), None) < 0
    assert b"SetTrajectory" in L.CassieGetLastError()
    assert L.Cassie2dBatchEnvStep(b.h, 0, 2, a.data_ptr(), 10, 0, obs.data_ptr(), r.data_ptr(), d.data_ptr(), None) < 0
    assert L.Cassie2dBatchSquat(b.h, 0, 1, None, None, None) < 0
    assert not L.Cassie2dBatchInit(0, 0, None, 32) and not L.Cassie2dBatchInit(4, 0, None, 16)
    assert not L.Cassie2dBatchInit(4, 0, b"/nonexistent/model.xml", 32) and b"cannot open" in L.CassieGetLastError()
    # the handle is still usable after errors
    b.step_torque(torch.zeros(8, 6), 1)
    assert torch.isfinite(b.get_general_state()).all()
    b.close()


def test_imitation_env_live_qstate(E, LIB, oracle, omodel):
    """cassie2d.py reward with the fix of SURVEY App. D.4 (reference_faithful=False): qstate follows the
    robot, so episodes no longer end at their first step; checked against the Python formula at EVERY policy
    step, teacher-forced (the PD-held robot amplifies rounding chaotically, DESIGN section 7, so each policy step
    starts from the oracle's state on both sides)."""
    from cassierl_b200.trajectory import Cassie2dTraj
    tr = Cassie2dTraj()
    n, T = 4, 8
    env = E.Cassie2dBatchEnv(n, task="imitate", control_mode="PD", precision=64, auto_reset=False, reference_faithful=False)
    env.reset()
    refs = [oracle.Cassie2d(omodel) for _ in range(n)]
    st = s26(oracle, QPOS_INIT_PY, np.zeros(13))
    for c in refs:
        c.reset(st); c.step_torque(np.zeros(6)); c.reset(st)   # fresh op-space state at reset (FRESH_OBS_ON_RESET)
    rng = np.random.default_rng(5)
    t = 0.0
    for k in range(T):
        if k > 0:
            S = np.array([s26(oracle, *c.data.state()) for c in refs]); Wm = np.array([c.data.warmstart() for c in refs])
            env.batch.reset(torch.tensor(S, dtype=torch.float64, device=env.batch.device))
            env.batch.set_warm_start(torch.tensor(Wm))
        A = tr.qpos[min(10 * (k + 1), 1681)][[3, 4, 6, 8, 9, 11]] + 0.02 * rng.standard_normal((n, 6))
        obs, rew, done = env.step(torch.tensor(A), n=10)
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        for _ in range(10):
            t += 0.0005
        ref = tr.state(t)[0][[0, 1, 2, 3, 4, 6, 8, 9, 11]]
        for e, c in enumerate(refs):
            for _ in range(10):
                c.step_pd(A[e])
            s = c.op_state(); q, _ = c.data.state()
            j = q[[3, 4, 6, 8, 9, 11]].sum() - ref[3:].sum()
            p = s[0] + s[1] - ref[0] - ref[1]; o = s[2] - ref[2]
            r = 0.5 * np.exp(-j * j) + 0.3 * np.exp(-p * p) + 0.1 * np.exp(-o * o)
            assert abs(rew[e] - r) < 1e-8, (k, e)
            assert np.array_equal(obs[e, 17:], ref), (k, e)
    env.terminate()


def test_auto_reset_returns_reset_observation(E, LIB, oracle):
    """with CASSIE_AUTO_RESET the observation of a done env is the one env.reset() returns (the caller's
    next action belongs to the new episode); CASSIE_TERMINAL_OBS keeps the terminal observation."""
    n = 64
    rng = np.random.default_rng(9)
    envs = {k: E.Cassie2dBatchEnv(n, task="stand", control_mode="Torque", precision=64, auto_reset=True, terminal_obs=k)
            for k in (False, True)}
    for env in envs.values():
        env.reset()
    seen = 0
    for k in range(80):
        a = torch.tensor(rng.uniform(-1, 1, (n, 6)) * TORQUE_HIGH)
        o0, r0, d0 = [x.clone() for x in envs[False].step(a)]
        o1, r1, d1 = [x.clone() for x in envs[True].step(a)]
        assert torch.equal(d0, d1) and torch.equal(r0, r1)
        dm = d0.bool()
        assert torch.equal(o0[~dm], o1[~dm])
        if dm.any():
            seen += int(dm.sum())
            assert (o1[dm][:, 0] < 0.5).all()                    # terminal observation: the fallen robot (z < 0.5)
            # reset observation: stale lagged op-space state (App. D.2) but the FRESH pitch / pitch rate of the reset pose
            assert (o0[dm][:, 1] == 0).all() and (o0[dm][:, 4] == 0).all()
            s = envs[False].batch.get_general_state()
            assert torch.allclose(s[dm][:, 1], torch.tensor(0.939, dtype=s.dtype, device=s.device))
            # the same thing EnvReset would now return for these envs: obs built from the stored op-space state
            o18 = envs[False].batch.get_operational_space_state()[dm]
            want = o18[:, 1:18].clone(); want[:, 5] -= o18[:, 0]; want[:, 11] -= o18[:, 0]
            assert torch.equal(o0[dm], want)
    assert seen > 0
    for env in envs.values():
        env.terminate()


def test_non_finite_env_is_reset_and_flagged(E, LIB, oracle):
    """a NaN / Inf env reports done, reward 0, a finite observation, is reset (solver state too) and
    flagged with status 3 in the stats; its neighbours are untouched (mj_checkPos/Vel/Acc [EXT])."""
    n = 32
    ref = E.Cassie2dBatchEnv(n, task="stand", control_mode="PD", precision=32, auto_reset=False)
    env = E.Cassie2dBatchEnv(n, task="stand", control_mode="PD", precision=32, auto_reset=False)
    ref.reset(); env.reset()
    a = torch.tensor(np.tile(QPOS_INIT_PY[[3, 4, 6, 8, 9, 11]], (n, 1)))
    s = env.batch.get_general_state().clone()
    s[5, 7] = float("nan"); s[9, 1] = float("inf"); s[20, 3] = 1e12
    env.batch.reset(s)
    o, r, d = [x.clone() for x in env.step(a)]
    o2, r2, d2 = [x.clone() for x in ref.step(a)]
    bad = torch.zeros(n, dtype=torch.bool, device=o.device); bad[[5, 9, 20]] = True
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
    assert (d[bad] == 1).all() and (r[bad] == 0).all()
    assert torch.equal(o[~bad], o2[~bad]) and torch.equal(r[~bad], r2[~bad]) and torch.equal(d[~bad], d2[~bad])
    st = env.batch.stats()
    assert (st[bad][:, 3] == LIB.STATUS_DIVERGED).all() and (st[~bad][:, 3] != LIB.STATUS_DIVERGED).all()
    g = env.batch.get_general_state()
    assert torch.isfinite(g).all() and torch.allclose(g[bad][:, 1], torch.tensor(0.939, device=g.device))
    assert (env.batch.get_warm_start()[bad] == 0).all()
    # and the env lives on
    o, r, d = env.step(a)
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
    env.terminate(); ref.terminate()


def test_imitation_reset_observation_has_zero_reference_slots(E, LIB):
    """env.reset() of cassie2d.py returns zeros in obs[17:26] (cassie2d.py:78-95); step() fills them."""
    env = E.Cassie2dBatchEnv(4, task="imitate", control_mode="PD", precision=64, auto_reset=False)
    o = env.reset().clone()
    assert (o[:, 17:] == 0).all() and (o[:, 0] > 0.9).all()
    a = torch.tensor(np.tile(QPOS_INIT_PY[[3, 4, 6, 8, 9, 11]], (4, 1)))
    o, _, _ = env.step(a)
    assert (o[:, 17:] != 0).any()
    env.terminate()


def philox_first(c0, c1, c2, k0, k1):
    """numpy restatement of csrc/cassie2d_api.cu philox_first (Philox4x32-10, counter (c0, c1, c2, 'SAMP'))"""
    c = [np.uint64(c0), np.uint64(c1), np.uint64(c2), np.uint64(0x53414D50)]
    k0 = np.uint64(k0); k1 = np.uint64(k1); M = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]; p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & M, p1 & M, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & M, p0 & M]
        k0 = (k0 + np.uint64(0x9E3779B9)) & M; k1 = (k1 + np.uint64(0xBB67AE85)) & M
    return int(c[0])


def test_random_phase_reset_on_device():
    """Cassie2dTraj.sample() on the device (cassie2d_trajectory.py:26-28): the sampled rows follow the documented Philox
    stream, every env restarts exactly at its row (qpos, qvel, env clock), the draw is invariant to the sharding, the
    mask leaves the other envs alone, and the first imitation reward uses the clock of the sampled row."""
    import torch
    from cassierl_b200 import envs
    from cassierl_b200.trajectory import Cassie2dTraj
    tr = Cassie2dTraj()
    rows = len(tr.time)
    n = 257
    env = envs.Cassie2dBatchEnv(n, task="imitate", control_mode="PD", precision=64)
    env.reset()
    obs, idx = env.reset_sampled(seed=77, draw=3, first_global_env=1000, want_index=True)
    idx = idx.cpu().numpy()
    want = np.array([philox_first(1000 + e, 3, 0, 77, 0) % rows for e in range(n)])
    assert (idx == want).all()
    assert len(np.unique(idx)) > n // 2 and idx.min() >= 0 and idx.max() < rows
    s = env.batch.get_general_state().cpu().numpy()
    q = np.concatenate([s[:, 0:3], s[:, 6:11], s[:, 16:21]], axis=1); v = np.concatenate([s[:, 3:6], s[:, 11:16], s[:, 21:26]], axis=1)
    assert np.abs(q - tr.qpos[idx]).max() == 0 and np.abs(v - tr.qvel[idx]).max() == 0
    o = obs.cpu().numpy()
    assert np.isfinite(o).all() and (o[:, 17:] == 0).all()            # reset(): reference slots are zero
    assert np.abs(o[:, 0] - tr.qpos[idx, 1]).max() < 1e-12              # obs[0] = pelvis z
    # sharding invariance: a sub-batch with the matching global offset draws the same rows
    env2 = envs.Cassie2dBatchEnv(64, task="imitate", control_mode="PD", precision=64)
    env2.reset()
    _, idx2 = env2.reset_sampled(seed=77, draw=3, first_global_env=1000 + 100, want_index=True)
    assert (idx2.cpu().numpy() == idx[100:164]).all()
    # mask: only the flagged envs move
    mask = torch.zeros(n, dtype=torch.uint8); mask[::2] = 1
    before = env.batch.get_general_state().clone()
    _, idx3 = env.reset_sampled(seed=78, draw=0, first_global_env=1000, mask=mask, want_index=True)
    after = env.batch.get_general_state()
    assert torch.equal(after[1::2], before[1::2]) and not torch.equal(after[::2], before[::2])
    # the env clock is the sampled row's time: one policy step later the reference row is index(time[i] + 10 dt)
    a = torch.tensor(np.tile(tr.qpos[0, [3, 4, 6, 8, 9, 11]], (64, 1)), dtype=torch.float64, device="cuda")
    o2, _, _ = env2.step(a, n=10)
    t1 = tr.time[idx2.cpu().numpy()] + 10 * 0.0005
    ref_rows = np.array([tr.index(t) for t in t1])
    near = np.array([min(abs(int(r) - int(tr.index(t - 1e-9))), abs(int(r) - int(tr.index(t + 1e-9)))) for r, t in zip(ref_rows, t1)])
    got = o2.cpu().numpy()[:, 17:]
    exp = tr.qpos[ref_rows][:, [0, 1, 2, 3, 4, 6, 8, 9, 11]]
    ok = np.abs(got - exp).max(axis=1) < 1e-12
    assert ok[near == 0].all() and ok.mean() > 0.9        # rows whose index sits on a floating-point boundary may differ by one
    env.terminate(); env2.terminate()


@pytest.mark.parametrize("mode", ["torque", "jacobian"])
def test_joint_limit_tiers_in_flight(E, LIB, oracle, omodel, mode):
    """Robots in the air with 3 or 4 joint limits on a leg next to standing ones (what random OSC accelerations produce,
    rllab/envs/cassie_stand2d.py:50).  Quad engine: the 20-row tier, voted per warp (torque: one warp per CTA) or per CTA
    (jacobian: seven lock-step warps); thread engine: its middle tier.  Every env matches the oracle facade to the fp64
    bar, and a standing env's result does not depend on which tier its CTA mates drag it through (bit for bit)."""
    from test_quad_host import flight_state
    n, steps = 120, 5
    S = np.zeros((n, 26))
    for e in range(n):
        if e % 3 == 0:
            q, v = QPOS_INIT_CTOR.copy(), np.zeros(13)
        else:
            q, v = flight_state(omodel, 7000 + e, 3 + e % 2, 4 - (e // 2) % 3)
        S[e] = s26(oracle, q, v)
    act = torch.zeros(n, 6, dtype=torch.float64)

    def run(states):
        b = E.Cassie2dBatch(n, precision=64)
        b.reset(torch.tensor(states))
        (b.step_torque if mode == "torque" else b.step_jacobian)(act, steps)
        out = b.get_general_state().cpu().numpy()
        rows = b.stats().cpu().numpy()[:, 0]
        b.close()
        return out, rows

    got, rows = run(S)
    assert rows.max() >= 4 + 3 + 2           # the limit rows were really there
    for e in range(0, n, 5):
        c = oracle.Cassie2d(omodel)
        c.reset(S[e])
        for _ in range(steps):
            (c.step_torque if mode == "torque" else c.step_jacobian)(np.zeros(6))
        assert rel_err(got[e], c.general_state()) < 1e-9, e
    alone, _ = run(np.repeat(S[0:1], n, axis=0))
    assert np.array_equal(alone[0], got[0])
