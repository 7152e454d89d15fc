"""TEST INFRASTRUCTURE -- ctypes wrapper of the fp64 CPU oracle (oracle/cassie_oracle.h).

PARITY UNPINNED (see the header of cassie_oracle.h).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product
package `cassierl_b200` never does.
"""
import ctypes as ct
import os
import subprocess
import numpy as np

from . import mjcf_reader

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcassie_oracle.so")
_lib = None

c_dp = ct.POINTER(ct.c_double)
c_ip = ct.POINTER(ct.c_int)


def build(force=False):
    """Compile the C restatement with gcc (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("cassie_oracle.c", "cassie_oracle_ctrl.c", "cassie_oracle.h", "cassie_oracle_internal.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _dp(a):
    return a.ctypes.data_as(c_dp)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = ct.CDLL(_LIB_PATH)
    vp = ct.c_void_p
    L.orc_model_new.restype = vp
    L.orc_model_rbdl_variant.restype = vp
    L.orc_model_rbdl_variant.argtypes = [vp]
    L.orc_data_new.restype = vp
    L.orc_data_new.argtypes = [vp]
    L.orc_kin_new.restype = vp
    L.orc_cassie_new.restype = vp
    L.orc_cassie_new.argtypes = [vp, vp]
    L.orc_cassie_data.restype = vp
    L.orc_cassie_data.argtypes = [vp]
    L.orc_total_mass.restype = ct.c_double
    L.orc_total_mass.argtypes = [vp]
    L.orc_get_time.restype = ct.c_double
    L.orc_get_time.argtypes = [vp]
    L.orc_energy.restype = ct.c_double
    L.orc_energy.argtypes = [vp, vp, c_dp, c_dp]
    L.orc_contact_mask.restype = ct.c_ulonglong
    L.orc_contact_mask.argtypes = [vp, vp]
    L.orc_rollout.restype = ct.c_long
    L.orc_rollout.argtypes = [vp, vp, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_dp, ct.c_int,
                              c_dp, c_dp, c_dp, ct.c_int]
    L.orc_qp_solve.restype = ct.c_int
    L.orc_pool_new.restype = vp
    L.orc_pool_new.argtypes = [vp, vp, ct.c_int]
    L.orc_pool_free.argtypes = [vp]
    L.orc_pool_run.restype = ct.c_long
    L.orc_pool_run.argtypes = [vp, ct.c_int, ct.c_int, ct.c_int, c_dp, ct.c_int, c_dp, c_dp, ct.c_int]
    L.orc_add_body.argtypes = [vp, ct.c_int, c_dp, c_dp, c_dp, ct.c_double, c_dp]
    L.orc_add_joint.argtypes = [vp, ct.c_int, ct.c_int, c_dp, c_dp, ct.c_double, ct.c_int, c_dp,
                                ct.c_double, ct.c_double, c_dp, c_dp]
    L.orc_add_geom.argtypes = [vp, ct.c_int, ct.c_int, c_dp, c_dp, c_dp, ct.c_int, ct.c_int,
                               ct.c_int, c_dp, c_dp, c_dp, ct.c_double, ct.c_double]
    L.orc_add_site.argtypes = [vp, ct.c_int, c_dp]
    L.orc_add_connect.argtypes = [vp, ct.c_int, ct.c_int, c_dp, c_dp, c_dp]
    L.orc_add_motor.argtypes = [vp, ct.c_int, ct.c_double, ct.c_int, c_dp]
    L.orc_set_option.argtypes = [vp, ct.c_double, ct.c_int, ct.c_double, ct.c_double, c_dp]
    L.orc_compile.argtypes = [vp]
    L.orc_add_joint.restype = ct.c_int
    L.orc_rollout_tree.restype = ct.c_long
    L.orc_rollout_tree.argtypes = [vp, ct.c_int, ct.c_int, ct.c_int, c_dp, c_dp, c_dp, ct.c_double, c_dp, c_dp,
                                   c_dp, c_dp, c_ip, ct.c_int]
    for name in ("orc_nv", "orc_nq", "orc_nbody"):
        getattr(L, name).argtypes = [vp]
    L.orc_get_consts.argtypes = [vp, c_dp, c_dp, c_dp, c_dp]
    L.orc_set_state.argtypes = [vp, c_dp, c_dp]
    L.orc_get_state.argtypes = [vp, c_dp, c_dp]
    L.orc_set_warmstart.argtypes = [vp, c_dp]
    L.orc_get_warmstart.argtypes = [vp, c_dp]
    L.orc_set_time.argtypes = [vp, ct.c_double]
    L.orc_forward.argtypes = [vp, vp, c_dp]
    L.orc_step.argtypes = [vp, vp, c_dp]
    L.orc_get_M.argtypes = [vp, c_dp]
    L.orc_get_vectors.argtypes = [vp, c_dp, c_dp, c_dp, c_dp, c_dp]
    for name in ("orc_get_nefc", "orc_get_ncon", "orc_get_solver_iter"):
        getattr(L, name).argtypes = [vp]
    L.orc_get_efc.argtypes = [vp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, c_ip]
    L.orc_get_contacts.argtypes = [vp, c_dp, c_dp, c_dp, c_ip]
    L.orc_body_pose.argtypes = [vp, ct.c_int, c_dp, c_dp]
    L.orc_site_pos.argtypes = [vp, vp, ct.c_int, c_dp]
    L.orc_kin_update.argtypes = [vp, vp, c_dp, c_dp]
    L.orc_kin_mass_matrix.argtypes = [vp, vp, c_dp]
    L.orc_kin_nonlinear_effects.argtypes = [vp, vp, c_dp]
    L.orc_kin_point_jacobian.argtypes = [vp, vp, ct.c_int, c_dp, c_dp, c_dp]
    L.orc_kin_point_pos_vel_acc.argtypes = [vp, vp, ct.c_int, c_dp, c_dp, c_dp, c_dp]
    L.orc_pinv.argtypes = [ct.c_int, ct.c_int, c_dp, ct.c_double, c_dp, c_dp]
    L.orc_qp_solve.argtypes = [ct.c_int, ct.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    for name in ("orc_cassie_reset", "orc_cassie_step_torque", "orc_cassie_step_pd",
                 "orc_cassie_step_jacobian", "orc_cassie_step_osc", "orc_cassie_get_general_state",
                 "orc_cassie_get_op_state", "orc_cassie_last_ctrl"):
        getattr(L, name).argtypes = [vp, c_dp]
    L.orc_cassie_dynamic_state.argtypes = [vp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.orc_cassie_osc_last.argtypes = [vp, c_dp, c_dp, c_ip]
    for name in ("orc_model_free", "orc_data_free", "orc_kin_free", "orc_cassie_free"):
        getattr(L, name).argtypes = [vp]
    _lib = L
    return L


def _solimp5(s):
    s = np.asarray(s, dtype=np.float64)
    if s.size == 3:  # MuJoCo 1.50 3-parameter form -> 2.x defaults midpoint .5, power 2
        s = np.concatenate([s, [0.5, 2.0]])
    return np.ascontiguousarray(s)


def default_model_path():
    return os.path.join(_HERE, "..", "cassierl_b200", "model", "cassie2d_stiff.xml")


class Model:
    """Compiled oracle model (+ its RBDL-style controller variant)."""

    def __init__(self, path=None):
        L = lib()
        self.spec = mjcf_reader.read_mjcf(path or default_model_path())
        sp = self.spec
        m = L.orc_model_new()
        c = np.ascontiguousarray
        for b in sp["bodies"][1:]:
            L.orc_add_body(m, b["parent"], _dp(c(b["pos"])), _dp(c(b["mat"].reshape(-1))),
                           _dp(c(b["ipos"])), b["mass"], _dp(c(b["inertia"])))
        self.joint_dof = []   # first dof of every MJCF joint (a free joint owns six)
        for j in sp["joints"]:
            self.joint_dof.append(L.orc_add_joint(
                m, j["body"], j["type"], _dp(c(j["axis"])), _dp(c(j["pos"])), j["ref"],
                int(j["limited"]), _dp(c(j["range"])), j["damping"], j["armature"],
                _dp(c(j["solref"])), _dp(_solimp5(j["solimp"]))))
        for g in sp["geoms"]:
            L.orc_add_geom(m, g["body"], g["type"], _dp(c(g["pos"])), _dp(c(g["mat"].reshape(-1))),
                           _dp(c(g["size"])), g["contype"], g["conaffinity"], g["condim"],
                           _dp(c(g["friction"])), _dp(c(g["solref"])), _dp(_solimp5(g["solimp"])),
                           g["margin"], g["gap"])
        for s in sp["sites"]:
            L.orc_add_site(m, s["body"], _dp(c(s["pos"])))
        for e in sp["equalities"]:
            L.orc_add_connect(m, e["body1"], e["body2"], _dp(c(e["anchor"])), _dp(c(e["solref"])),
                              _dp(_solimp5(e["solimp"])))
        for a in sp["actuators"]:
            L.orc_add_motor(m, self.joint_dof[a["joint"]], a["gear"], int(a["ctrllimited"]), _dp(c(a["ctrlrange"])))
        o = sp["option"]
        L.orc_set_option(m, o["timestep"], o["iterations"], o["tolerance"], o["impratio"],
                         _dp(c(o["gravity"])))
        if L.orc_compile(m) != 0:
            raise RuntimeError("oracle: model compile failed")
        self.ptr = m
        self.rbdl_ptr = L.orc_model_rbdl_variant(m)
        self.nv = L.orc_nv(m)
        self.nq = L.orc_nq(m)
        self.nbody = L.orc_nbody(m)
        self.nu = len(sp["actuators"])

    def total_mass(self):
        return lib().orc_total_mass(self.ptr)

    def consts(self, rbdl=False):
        L = lib()
        p = self.rbdl_ptr if rbdl else self.ptr
        biw = np.zeros((self.nbody, 2)); diw = np.zeros(self.nv); mi = ct.c_double(0)
        a2 = np.zeros((len(self.spec["equalities"]), 3))
        L.orc_get_consts(p, _dp(biw), _dp(diw), ct.byref(mi), _dp(a2))
        return dict(body_invweight0=biw, dof_invweight0=diw, meaninertia=mi.value, eq_anchor2=a2)


class Data:
    """mjData equivalent."""

    def __init__(self, model, ptr=None):
        self.m = model
        self.ptr = ptr if ptr is not None else lib().orc_data_new(model.ptr)
        self._own = ptr is None

    def set_state(self, qpos, qvel):
        lib().orc_set_state(self.ptr, _dp(np.ascontiguousarray(qpos, np.float64)),
                            _dp(np.ascontiguousarray(qvel, np.float64)))

    def state(self):
        q = np.zeros(self.m.nq); v = np.zeros(self.m.nv)
        lib().orc_get_state(self.ptr, _dp(q), _dp(v))
        return q, v

    def warmstart(self):
        w = np.zeros(self.m.nv)
        lib().orc_get_warmstart(self.ptr, _dp(w))
        return w

    def set_warmstart(self, w):
        lib().orc_set_warmstart(self.ptr, _dp(np.ascontiguousarray(w, np.float64)))

    @property
    def time(self):
        return lib().orc_get_time(self.ptr)

    def forward(self, ctrl=None):
        c = None if ctrl is None else _dp(np.ascontiguousarray(ctrl, np.float64))
        lib().orc_forward(self.m.ptr, self.ptr, c)

    def step(self, ctrl=None):
        c = None if ctrl is None else _dp(np.ascontiguousarray(ctrl, np.float64))
        lib().orc_step(self.m.ptr, self.ptr, c)

    def M(self):
        M = np.zeros((self.m.nv, self.m.nv))
        lib().orc_get_M(self.ptr, _dp(M))
        return M

    def vectors(self):
        n = self.m.nv
        out = [np.zeros(n) for _ in range(5)]
        lib().orc_get_vectors(self.ptr, *[_dp(o) for o in out])
        return dict(zip(("qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "qacc"), out))

    def efc(self):
        L = lib()
        n = L.orc_get_nefc(self.ptr)
        nv = self.m.nv
        J = np.zeros((n, nv)); pos = np.zeros(n); aref = np.zeros(n); R = np.zeros(n); f = np.zeros(n)
        tp = np.zeros(n, np.int32); idd = np.zeros(n, np.int32)
        L.orc_get_efc(self.ptr, _dp(J), _dp(pos), _dp(aref), _dp(R), _dp(f),
                      tp.ctypes.data_as(c_ip), idd.ctypes.data_as(c_ip))
        return dict(J=J, pos=pos, aref=aref, R=R, force=f, type=tp, id=idd,
                    iters=L.orc_get_solver_iter(self.ptr))

    def contacts(self):
        L = lib()
        n = L.orc_get_ncon(self.ptr)
        dist = np.zeros(n); pos = np.zeros((n, 3)); frame = np.zeros((n, 3, 3)); g = np.zeros(n, np.int32)
        L.orc_get_contacts(self.ptr, _dp(dist), _dp(pos), _dp(frame), g.ctypes.data_as(c_ip))
        return dict(dist=dist, pos=pos, frame=frame, geom=g)

    def contact_mask(self):
        return int(lib().orc_contact_mask(self.m.ptr, self.ptr))

    def body_pose(self, b):
        p = np.zeros(3); R = np.zeros((3, 3))
        lib().orc_body_pose(self.ptr, b, _dp(p), _dp(R))
        return p, R

    def site_pos(self, s):
        p = np.zeros(3)
        lib().orc_site_pos(self.m.ptr, self.ptr, s, _dp(p))
        return p

    def energy(self):
        T = ct.c_double(0); U = ct.c_double(0)
        E = lib().orc_energy(self.m.ptr, self.ptr, ct.byref(T), ct.byref(U))
        return E, T.value, U.value


class Kin:
    """RBDL-equivalent kinematics at an arbitrary (q, qd) on the physics or controller model."""

    def __init__(self, model, rbdl=True):
        self.m = model
        self.mp = model.rbdl_ptr if rbdl else model.ptr
        self.ptr = lib().orc_kin_new()

    def update(self, q, qd):
        lib().orc_kin_update(self.mp, self.ptr, _dp(np.ascontiguousarray(q, np.float64)),
                             _dp(np.ascontiguousarray(qd, np.float64)))

    def mass_matrix(self):
        M = np.zeros((self.m.nv, self.m.nv))
        lib().orc_kin_mass_matrix(self.mp, self.ptr, _dp(M))
        return M

    def nonlinear_effects(self):
        b = np.zeros(self.m.nv)
        lib().orc_kin_nonlinear_effects(self.mp, self.ptr, _dp(b))
        return b

    def point_jacobian(self, body, p):
        jp = np.zeros((3, self.m.nv)); jr = np.zeros((3, self.m.nv))
        lib().orc_kin_point_jacobian(self.mp, self.ptr, body, _dp(np.ascontiguousarray(p, np.float64)),
                                     _dp(jp), _dp(jr))
        return jp, jr

    def point(self, body, p):
        pos = np.zeros(3); vel = np.zeros(3); acc = np.zeros(3)
        lib().orc_kin_point_pos_vel_acc(self.mp, self.ptr, body,
                                        _dp(np.ascontiguousarray(p, np.float64)), _dp(pos), _dp(vel), _dp(acc))
        return pos, vel, acc


def pinv(A, tol):
    A = np.ascontiguousarray(A, np.float64)
    r, c = A.shape
    out = np.zeros((c, r)); sv = np.zeros(min(r, c))
    lib().orc_pinv(r, c, _dp(A), tol, _dp(out), _dp(sv))
    return out, sv


def qp_solve(G, g, A, lbA, ubA, lb, ub, x0):
    n = G.shape[0]
    mc = 0 if A is None else A.shape[0]
    x = np.ascontiguousarray(x0, np.float64).copy()
    f = lambda a: None if a is None else _dp(np.ascontiguousarray(a, np.float64))
    keep = [np.ascontiguousarray(a, np.float64) if a is not None else None for a in (G, g, A, lbA, ubA, lb, ub)]
    it = lib().orc_qp_solve(n, mc, *[None if k is None else _dp(k) for k in keep], _dp(x))
    return x, it


class Cassie2d:
    """The reference's Cassie2d facade (Cassie2d.cpp:29-237) on the oracle."""

    QPOS_INIT_PY = np.array([0.0, 0.939, 0.0, 0.68111815, -1.40730357, 1.62972042, -1.77611107,
                             -0.61968407, 0.68111815, -1.40730357, 1.62972042, -1.77611107, -0.61968407])

    def __init__(self, model):
        self.m = model
        self.ptr = lib().orc_cassie_new(model.ptr, model.rbdl_ptr)
        self.data = Data(model, ptr=lib().orc_cassie_data(self.ptr))

    def _call(self, name, arr, n):
        a = np.ascontiguousarray(arr, np.float64)
        assert a.size == n
        getattr(lib(), name)(self.ptr, _dp(a))

    def reset(self, state26):
        self._call("orc_cassie_reset", state26, 26)

    def step_torque(self, u):
        self._call("orc_cassie_step_torque", u, 6)

    def step_pd(self, a):
        self._call("orc_cassie_step_pd", a, 6)

    def step_jacobian(self, f):
        self._call("orc_cassie_step_jacobian", f, 6)

    def step_osc(self, a):
        self._call("orc_cassie_step_osc", a, 7)

    def general_state(self):
        s = np.zeros(26)
        lib().orc_cassie_get_general_state(self.ptr, _dp(s))
        return s

    def op_state(self):
        s = np.zeros(18)
        lib().orc_cassie_get_op_state(self.ptr, _dp(s))
        return s

    def last_ctrl(self):
        u = np.zeros(6)
        lib().orc_cassie_last_ctrl(self.ptr, _dp(u))
        return u

    def dynamic_state(self):
        M = np.zeros((13, 13)); b = np.zeros(13); Bt = np.zeros((13, 6)); Jc = np.zeros((12, 13))
        Jeq = np.zeros((6, 13)); jd = np.zeros(6)
        lib().orc_cassie_dynamic_state(self.ptr, _dp(M), _dp(b), _dp(Bt), _dp(Jc), _dp(Jeq), _dp(jd))
        return dict(M=M, bias=b, Bt=Bt, Jc=Jc, Jeq=Jeq, JeqdotQdot=jd)

    def osc_last(self):
        x = np.zeros(39); obj = ct.c_double(0); it = ct.c_int(0)
        lib().orc_cassie_osc_last(self.ptr, _dp(x), ct.byref(obj), ct.byref(it))
        return x, obj.value, it.value


def state26_from_qpos_qvel(qpos, qvel):
    """StateGeneral memory order (RobotInterface.h:38-45, cassie2d_structs.py:77-98)."""
    s = np.zeros(26)
    s[0:3] = qpos[0:3]; s[3:6] = qvel[0:3]
    s[6:11] = qpos[3:8]; s[11:16] = qvel[3:8]
    s[16:21] = qpos[8:13]; s[21:26] = qvel[8:13]
    return s


def rollout(model, n_envs, n_steps, mode, actions=None, hold=1, phase=None, init26=None, n_threads=0):
    out = np.zeros((n_envs, 26))
    adim = 0 if actions is None else actions.shape[-1]
    a = None if actions is None else np.ascontiguousarray(actions, np.float64)
    ph = None if phase is None else np.ascontiguousarray(phase, np.float64)
    i26 = None if init26 is None else np.ascontiguousarray(init26, np.float64)
    n = lib().orc_rollout(model.ptr, model.rbdl_ptr, n_envs, n_steps, mode, hold,
                          None if a is None else _dp(a), adim, None if ph is None else _dp(ph),
                          None if i26 is None else _dp(i26), _dp(out), n_threads)
    return n, out


def model3d_path():
    return os.path.join(_HERE, "..", "cassierl_b200", "model", "cassie3d_stiff.xml")


def rollout_tree(model, qpos0, qvel0, n_steps, actions=None, hold=10, z_done=0.0, reset_qpos=None, reset_qvel=None,
                 n_threads=0):
    """3-D torque rollouts (orc_rollout_tree): qpos0 [n, nq], qvel0 [n, nv], actions [n, ceil(n_steps/hold), nu]."""
    q0 = np.ascontiguousarray(qpos0, np.float64); v0 = np.ascontiguousarray(qvel0, np.float64)
    n = q0.shape[0]
    oq = np.zeros((n, model.nq)); ov = np.zeros((n, model.nv)); rs = np.zeros(n, np.int32)
    a = None if actions is None else np.ascontiguousarray(actions, np.float64)
    rq = None if reset_qpos is None else np.ascontiguousarray(reset_qpos, np.float64)
    rv = None if reset_qvel is None else np.ascontiguousarray(reset_qvel, np.float64)
    tot = lib().orc_rollout_tree(model.ptr, n, n_steps, hold, None if a is None else _dp(a), _dp(q0), _dp(v0),
                                 float(z_done), None if rq is None else _dp(rq), None if rv is None else _dp(rv),
                                 _dp(oq), _dp(ov), rs.ctypes.data_as(c_ip), n_threads)
    return tot, oq, ov, rs


class Pool:
    """Persistent envs for bench.py --impl reference (facades, QP hot start and squat clocks survive across run() calls)."""

    def __init__(self, model, n_envs):
        self.m, self.n = model, n_envs
        self.ptr = lib().orc_pool_new(model.ptr, model.rbdl_ptr, n_envs)

    def run(self, n_steps, mode, actions=None, hold=1, phase=None, want_state=False, n_threads=0):
        out = np.zeros((self.n, 26)) if want_state else None
        adim = 0 if actions is None else actions.shape[-1]
        a = None if actions is None else np.ascontiguousarray(actions, np.float64)
        ph = None if phase is None else np.ascontiguousarray(phase, np.float64)
        n = lib().orc_pool_run(self.ptr, n_steps, mode, hold, None if a is None else _dp(a), adim,
                               None if ph is None else _dp(ph), None if out is None else _dp(out), n_threads)
        return n, out

    def close(self):
        if self.ptr:
            lib().orc_pool_free(self.ptr)
            self.ptr = None
