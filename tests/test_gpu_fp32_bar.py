"""The fp32 product build against the north-star bar, on the GPU, through the C-ABI.

Protocol ("identical initial states and action sequences", BASELINE.json north_star): a state of the oracle's
own trajectory is rounded to fp32 (qpos, qvel, solver warm start, action) and handed to BOTH sides, so the
number measures the device's arithmetic, not the rounding of its input.  One Step* call each; compared are
qpos/qvel after the step.

Error metrics (conftest): `rel_err` = max |a-b| / max(1, |b|), i.e. relative for |value| >= 1 and ABSOLUTE below
1 (most joint velocities and angles) -- that is the metric the 1e-5 bar is asserted on; `norm_rel_err` =
||a-b||_inf / ||b||_inf per state vector (qpos, qvel separately), the true norm-wise relative error, is asserted
at the stated looser bound 1e-4 and printed (qvel of a robot near rest is ~1e-2, so 1e-5 of its norm would be
1e-7 rad/s, below what one fp32 M^-1 application resolves).
"""
import numpy as np
import pytest

from conftest import QPOS_INIT_CTOR, norm_rel_err, rel_err, squat_jacobian_action, squat_osc_action

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from cassierl_b200 import envs
    return envs


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32).astype(np.float64)


def q_from_s26(s):
    s = np.asarray(s)
    q = np.concatenate([s[..., 0:3], s[..., 6:11], s[..., 16:21]], axis=-1)
    v = np.concatenate([s[..., 3:6], s[..., 11:16], s[..., 21:26]], axis=-1)
    return q, v


def oracle_step_from(oracle, omodel, mode, q, v, w, a):
    """one Step* of the oracle facade from an explicit (qpos, qvel, warm start)"""
    c = oracle.Cassie2d(omodel)
    c.reset(oracle.state26_from_qpos_qvel(q, v))
    c.data.set_warmstart(w)
    (c.step_torque, c.step_pd, c.step_jacobian, c.step_osc)[mode](a)
    q1, v1 = c.data.state()
    return np.concatenate([q1, v1]), c.data.contact_mask()


def collect(oracle, omodel, mode, action_fn, n_steps, pick, phase=0.0):
    """runs the oracle facade in closed loop and returns fp32-rounded (q, v, warm, action) at the picked steps"""
    c = oracle.Cassie2d(omodel)
    out = []
    for k in range(n_steps):
        a = action_fn(k, c, phase)
        if k in pick:
            q, v = c.data.state()
            out.append((f32(q), f32(v), f32(c.data.warmstart()), f32(a)))
        (c.step_torque, c.step_pd, c.step_jacobian, c.step_osc)[mode](a)
    return out


def run_gpu_single_steps(E, oracle, mode, samples):
    n = len(samples)
    b = E.Cassie2dBatch(n, precision=32)
    S = np.array([oracle.state26_from_qpos_qvel(s[0], s[1]) for s in samples])
    b.reset(torch.tensor(S, dtype=torch.float32, device=b.device))
    b.set_warm_start(torch.tensor(np.array([s[2] for s in samples])))
    mask = torch.zeros(n, dtype=torch.int32, device=b.device)
    b.step(mode, torch.tensor(np.array([s[3] for s in samples])), 1, contact_mask=mask)
    q, v = q_from_s26(b.get_general_state().cpu().numpy().astype(np.float64))
    st = b.stats().cpu().numpy()
    b.close()
    return q, v, mask.cpu().numpy().astype(np.uint64), st


def check(oracle, omodel, mode, samples, q, v, mask, tag, bar=1e-5):
    worst = 0.0; worst_n = 0.0
    for i, s in enumerate(samples):
        ref, rmask = oracle_step_from(oracle, omodel, mode, *s)
        e = rel_err(np.concatenate([q[i], v[i]]), ref)
        en = max(norm_rel_err(q[i], ref[:13]), norm_rel_err(v[i], ref[13:]))
        worst = max(worst, e); worst_n = max(worst_n, en)
        assert int(mask[i]) == int(rmask), (tag, i)
    print("%s: %d single steps, worst |a-b|/max(1,|b|) = %.2e, worst norm-wise relative = %.2e" % (tag, len(samples), worst, worst_n))
    assert worst < bar, (tag, worst)
    assert worst_n < 1e-4, (tag, worst_n)


def test_fp32_single_step_pd_gpu(E, oracle, omodel):
    """StepPd (Cassie2d.cpp:96-117) on the GPU in fp32: 48 states of a PD-held, toe-chattering robot."""
    rng = np.random.default_rng(7)
    tg = QPOS_INIT_CTOR[[3, 4, 6, 8, 9, 11]] + rng.uniform(-0.1, 0.1, (40, 6))
    samples = collect(oracle, omodel, 1, lambda k, c, ph: tg[k // 10], 400, set(range(3, 400, 8)))[:48]
    q, v, mask, _ = run_gpu_single_steps(E, oracle, 1, samples)
    check(oracle, omodel, 1, samples, q, v, mask, "fp32 PD")


def test_fp32_single_step_jacobian_gpu(E, oracle, omodel):
    """StepJacobian (Cassie2d.cpp:119-177) on the GPU in fp32 along the squatting.py stream, 4 phases."""
    samples = []
    for ph in (0.0, 1.6, 3.1, 4.7):
        samples += collect(oracle, omodel, 2, lambda k, c, p: squat_jacobian_action(c.op_state(), k * 0.0005, p), 300,
                           set(range(5, 300, 25)), ph)
    q, v, mask, _ = run_gpu_single_steps(E, oracle, 2, samples)
    check(oracle, omodel, 2, samples, q, v, mask, "fp32 Jacobian")


def test_fp32_single_step_osc_unloading_gpu(E, oracle, omodel):
    """StepOsc (Cassie2d.cpp:179-209) on the GPU in fp32 over the whole squat period INCLUDING the regime where the
    commanded descent unloads the feet and the QP sits at the apex of the friction pyramids (phases around pi,
    steps 20-200), which the benign-stream test does not visit."""
    samples = []
    for ph in np.linspace(0.0, 2 * np.pi, 16, endpoint=False):
        samples += collect(oracle, omodel, 3, lambda k, c, p: squat_osc_action(c.op_state(), k * 0.0005, p), 240,
                           set(range(20, 240, 20)), ph)
    q, v, mask, st = run_gpu_single_steps(E, oracle, 3, samples)
    assert (st[:, 3] == 0).all()
    check(oracle, omodel, 3, samples, q, v, mask, "fp32 OSC incl. unloading")


def test_config3_16384_fp32_osc_subset_vs_oracle(E, oracle, omodel):
    """BASELINE configs[2] at its real size: 16384 envs, fp32, OSC squatting controller in the kernel.  After 300
    closed-loop steps on the device, ONE more step of the same squat kernel is compared, for 24 envs spread over
    the batch, with the oracle started from the device's own state (qpos, qvel, warm start) and fed the action the
    squat law gives on the device's lagged op-space state and clock."""
    n, pre = 16384, 300
    subset = np.array([0, 1, 2, 3, 31, 32, 33, 255, 256, 1000, 2047, 2048, 4095, 4096, 5000, 8191, 8192, 9999, 12287,
                       12288, 14000, 16000, 16382, 16383])
    phase = 2 * np.pi * np.arange(n) / n
    b = E.Cassie2dBatch(n, precision=32)
    ph_d = torch.tensor(phase)
    b.squat(3, pre, phase=ph_d)
    s0 = b.get_general_state().cpu().numpy().astype(np.float64)[subset]
    w0 = b.get_warm_start().cpu().numpy().astype(np.float64)[subset]
    o0 = b.get_operational_space_state().cpu().numpy().astype(np.float64)[subset]
    mask = torch.zeros(n, dtype=torch.int32, device=b.device)
    b.squat(3, 1, phase=ph_d, contact_mask=mask)
    s1 = b.get_general_state()
    assert torch.isfinite(s1).all()
    st = b.stats().cpu().numpy()
    assert (st[:, 3] == 0).all()
    s1 = s1.cpu().numpy().astype(np.float64)[subset]
    m1 = mask.cpu().numpy().astype(np.uint64)[subset]
    b.close()
    t = 0.0
    for _ in range(pre):
        t = t + 0.0005                      # squatting.py:15, accumulated like the kernel's clock
    worst = 0.0
    for i, e in enumerate(subset):
        q, v = q_from_s26(s0[i])
        # the device evaluates the squat law in fp32 on its fp32 op-space state (targets in double)
        zt = np.float32(0.7 + 0.25 * np.sin(0.5 * 3.1415 * t + np.float64(np.float32(phase[e]))))
        zdt = np.float32(0.25 * np.cos(0.5 * 3.1415 * t + np.float64(np.float32(phase[e]))))
        o = o0[i].astype(np.float32)
        xt = (o[6] + o[12]) / np.float32(2)
        a = np.array([np.float32(100) * (xt - o[0]) + np.float32(20) * (np.float32(0) - o[3]),
                      np.float32(100) * (zt - o[1]) + np.float32(20) * (zdt - o[4]),
                      0.0, np.float32(100) * (np.float32(-5e-3) - o[7]), 0.0, np.float32(100) * (np.float32(-5e-3) - o[13]),
                      np.float32(20) * (np.float32(0) - o[2]) + np.float32(10) * (np.float32(0) - o[5])], np.float64)
        ref, rmask = oracle_step_from(oracle, omodel, 3, q, v, w0[i], a)
        q1, v1 = q_from_s26(s1[i])
        worst = max(worst, rel_err(np.concatenate([q1, v1]), ref))
        assert int(m1[i]) == int(rmask), e
    print("config 3 at 16384 envs: 24 envs, one squat-kernel step vs oracle, worst %.2e" % worst)
    assert worst < 1e-5, worst
