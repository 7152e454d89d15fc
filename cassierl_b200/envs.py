"""Host-side mirror of the reference's Python env interface over a batch dimension.

`Cassie2dBatch`  = the reference's ctypes surface (rllab/envs/cassie2d.py:22-50: Reset, StepOsc,
                   StepTorque, StepJacobian, StepPd, GetGeneralState, GetOperationalSpaceState)
                   for N envs living in HBM; arrays are torch CUDA tensors [N, dim].
`Cassie2dBatchEnv` = the rllab Env the reference builds on top of it (cassie_stand2d.py:56-275 /
                   cassie2d.py:57-372): reset() / step(action, n=10) -> obs, reward, done, with the
                   observation, reward and termination computed on device in the same launch.
PyTorch is used for device memory and streams only; all arithmetic is in libcassie2d.so.
"""
import numpy as np
import torch

from . import lib as _lib
from .trajectory import Cassie2dTraj

ACTION_DIM = {_lib.MODE_TORQUE: 6, _lib.MODE_PD: 6, _lib.MODE_JACOBIAN: 6, _lib.MODE_OSC: 7}
MODE_BY_NAME = {"Torque": _lib.MODE_TORQUE, "PD": _lib.MODE_PD, "Jacobian": _lib.MODE_JACOBIAN, "OSC": _lib.MODE_OSC}
OBS_DIM = {_lib.TASK_STAND: 17, _lib.TASK_IMITATE: 26}


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class Cassie2dBatch:
    """N Cassie2d instances stepped by one CUDA launch per call."""

    def __init__(self, n_envs, device=0, precision=32, xml_path=None):
        if not torch.cuda.is_available():
            raise RuntimeError("cassierl_b200 needs a CUDA device (there is no CPU fallback)")
        self.L = _lib.load()
        self.n = int(n_envs)
        self.device = torch.device("cuda", int(device))
        self.precision = int(precision)
        self.dtype = torch.float64 if precision == 64 else torch.float32
        torch.cuda.init()
        with torch.cuda.device(self.device):
            self.h = self.L.Cassie2dBatchInit(self.n, int(device), xml_path.encode() if xml_path else None, self.precision)
        if not self.h:
            raise RuntimeError("Cassie2dBatchInit failed: " + _lib.last_error())

    # -- helpers
    def _t(self, x, dim):
        t = torch.as_tensor(x, dtype=self.dtype, device=self.device)
        if t.dim() == 1:
            t = t.expand(self.n, dim)
        t = t.contiguous()
        assert t.shape == (self.n, dim), (t.shape, (self.n, dim))
        return t

    def empty(self, dim, dtype=None):
        shape = (self.n, dim) if dim else (self.n,)
        return torch.empty(shape, dtype=dtype or self.dtype, device=self.device)

    # -- the reference's FFI surface, batched
    def reset(self, state26=None, mask=None):
        """Reset (Cassie2d.cpp:78-82).  state26: 26 doubles (one StateGeneral) or tensor [N,26]."""
        with torch.cuda.device(self.device):
            if isinstance(state26, torch.Tensor) and state26.dim() == 2:
                assert mask is None
                s = self._t(state26, 26)
                _lib.check(self.L.Cassie2dBatchSetState(self.h, s.data_ptr(), _stream_ptr()), "SetState")
                return
            m = None
            if mask is not None:
                m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            sp = None
            if state26 is not None:
                arr = np.ascontiguousarray(np.asarray(state26, dtype=np.float64).reshape(26))
                sp = arr.ctypes.data_as(_lib.ct.POINTER(_lib.ct.c_double))
            _lib.check(self.L.Cassie2dBatchReset(self.h, m.data_ptr() if m is not None else None, sp, _stream_ptr()), "Reset")

    def step(self, mode, action, n_substeps=1, contact_mask=None):
        a = self._t(action, ACTION_DIM[mode])
        with torch.cuda.device(self.device):
            _lib.check(self.L.Cassie2dBatchStep(self.h, mode, a.data_ptr(), int(n_substeps),
                                                contact_mask.data_ptr() if contact_mask is not None else None,
                                                _stream_ptr()), "Step")

    def step_torque(self, torques, n_substeps=1, **kw):
        self.step(_lib.MODE_TORQUE, torques, n_substeps, **kw)

    def step_pd(self, angles, n_substeps=1, **kw):
        self.step(_lib.MODE_PD, angles, n_substeps, **kw)

    def step_jacobian(self, forces, n_substeps=1, **kw):
        self.step(_lib.MODE_JACOBIAN, forces, n_substeps, **kw)

    def step_osc(self, accels, n_substeps=1, **kw):
        self.step(_lib.MODE_OSC, accels, n_substeps, **kw)

    def get_general_state(self, out=None):
        out = self.empty(26) if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self.L.Cassie2dBatchGetGeneralState(self.h, out.data_ptr(), _stream_ptr()), "GetGeneralState")
        return out

    def get_operational_space_state(self, out=None):
        out = self.empty(18) if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self.L.Cassie2dBatchGetOperationalSpaceState(self.h, out.data_ptr(), _stream_ptr()), "GetOpState")
        return out

    def squat(self, mode, n_steps, phase=None, contact_mask=None):
        """squatting.py:8-16 on device (standing_controller_jacobian / _osc in the loop)."""
        p = None
        if phase is not None:
            p = torch.as_tensor(phase, dtype=self.dtype, device=self.device).contiguous()
            assert p.shape == (self.n,)
        with torch.cuda.device(self.device):
            _lib.check(self.L.Cassie2dBatchSquat(self.h, mode, int(n_steps), p.data_ptr() if p is not None else None,
                                                 contact_mask.data_ptr() if contact_mask is not None else None,
                                                 _stream_ptr()), "Squat")

    def set_warm_start(self, qacc):
        w = self._t(qacc, 13)
        with torch.cuda.device(self.device):
            _lib.check(self.L.Cassie2dBatchSetWarmStart(self.h, w.data_ptr(), _stream_ptr()), "SetWarmStart")

    def get_warm_start(self):
        out = self.empty(13)
        with torch.cuda.device(self.device):
            _lib.check(self.L.Cassie2dBatchGetWarmStart(self.h, out.data_ptr(), _stream_ptr()), "GetWarmStart")
        return out

    def stats(self):
        out = torch.empty((self.n, 4), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.Cassie2dBatchGetStats(self.h, out.data_ptr(), _stream_ptr()), "GetStats")
        return out

    # -- host-buffer (end-to-end) variants: numpy / pinned tensors in and out
    def step_host(self, mode, action_host, n_substeps, state_out_host=None):
        _lib.check(self.L.Cassie2dBatchStepHost(self.h, mode, action_host.data_ptr(), int(n_substeps),
                                                state_out_host.data_ptr() if state_out_host is not None else None), "StepHost")

    def squat_host(self, mode, n_steps, phase_host=None, state_out_host=None):
        _lib.check(self.L.Cassie2dBatchSquatHost(self.h, mode, int(n_steps),
                                                 phase_host.data_ptr() if phase_host is not None else None,
                                                 state_out_host.data_ptr() if state_out_host is not None else None), "SquatHost")

    def sync(self):
        _lib.check(self.L.Cassie2dBatchSync(self.h), "Sync")

    def close(self):
        if getattr(self, "h", None):
            self.L.Cassie2dBatchDestroy(self.h)
            self.h = None

    terminate = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Cassie2dBatchEnv:
    """Batched counterpart of the reference's `Cassie2dEnv` (rllab Env protocol).

    task='stand'   -> cassie_stand2d.py (17-d obs, alive-bonus reward, done when z < 0.5)
    task='imitate' -> cassie2d.py (26-d obs, reference-motion reward; with the default
                      reference_faithful=True it reproduces the reference's frozen-qstate reward).
    """

    def __init__(self, n_envs, device=0, task="stand", control_mode="OSC", precision=32, auto_reset=True,
                 reference_faithful=True, trajectory=None, terminal_obs=False):
        self.batch = Cassie2dBatch(n_envs, device, precision)
        self.n = self.batch.n
        self.task = _lib.TASK_STAND if task == "stand" else _lib.TASK_IMITATE
        self.mode = MODE_BY_NAME[control_mode]
        if self.mode == _lib.MODE_JACOBIAN:
            raise ValueError("the reference envs offer the OSC, Torque and PD action spaces")
        self.control_mode = control_mode
        self.flags = (_lib.AUTO_RESET if auto_reset else 0)
        if not reference_faithful:
            self.flags |= _lib.FRESH_OBS_ON_RESET | _lib.LIVE_QSTATE
        if terminal_obs:   # auto-reset returns the terminal observation instead of the new episode's first one
            self.flags |= _lib.TERMINAL_OBS
        if self.task == _lib.TASK_IMITATE:
            tr = trajectory if trajectory is not None else Cassie2dTraj()
            q = np.ascontiguousarray(tr.qpos, np.float64)
            _lib.check(self.batch.L.Cassie2dBatchSetTrajectory(
                self.batch.h, q.ctypes.data_as(_lib.ct.POINTER(_lib.ct.c_double)), q.shape[0], float(tr.time[-1])), "SetTrajectory")
        self.obs_dim = OBS_DIM[self.task]
        self.action_dim = ACTION_DIM[self.mode]
        b = self.batch
        self._obs = b.empty(self.obs_dim)
        self._rew = b.empty(0)
        self._done = b.empty(0, torch.uint8)

    # rllab spaces as (low, high) arrays: cassie_stand2d.py:240-275 / cassie2d.py:345-368
    @property
    def observation_space(self):
        high = np.full((self.obs_dim,), 1e20)
        return -high, high

    @property
    def action_space(self):
        if self.control_mode == "OSC":
            high = np.full((7,), 2e1)
            low = np.array([-2e1, -2e1, -2e1, 0, -2e1, 0, -2e1])
        elif self.control_mode == "Torque":
            high = np.array([12.0, 12.0, 0.9, 12.0, 12.0, 0.9])
            low = -1.0 * high
        else:
            high = np.radians([80.0, -37.0, -30.0, 80.0, -37.0, -30.0])
            low = np.radians([-50.0, -164.0, -140.0, -50.0, -164.0, -140.0])
        return low, high

    def reset(self):
        b = self.batch
        with torch.cuda.device(b.device):
            _lib.check(b.L.Cassie2dBatchEnvReset(b.h, self.task, self.flags, self._obs.data_ptr(), _stream_ptr()), "EnvReset")
        return self._obs

    def reset_sampled(self, seed=1, draw=0, mask=None, first_global_env=0, trajectory=None, want_index=False):
        """Random-phase reset: every env (or the masked ones) restarts at a random row of the reference trajectory --
        Cassie2dTraj.sample() (rllab/envs/cassie2d_trajectory.py:26-28) on the device, Philox-keyed by (seed, global env
        id, draw).  Returns the observation (and the sampled row indices with want_index)."""
        b = self.batch
        dp = _lib.ct.POINTER(_lib.ct.c_double)
        if not getattr(self, "_traj_detail", False):
            tr = trajectory if trajectory is not None else Cassie2dTraj()
            q = np.ascontiguousarray(tr.qpos, np.float64)
            v = np.ascontiguousarray(tr.qvel, np.float64); t = np.ascontiguousarray(tr.time, np.float64)
            if self.task != _lib.TASK_IMITATE:
                _lib.check(b.L.Cassie2dBatchSetTrajectory(b.h, q.ctypes.data_as(dp), q.shape[0], float(tr.time[-1])), "SetTrajectory")
            _lib.check(b.L.Cassie2dBatchSetTrajectoryDetail(b.h, v.ctypes.data_as(dp), t.ctypes.data_as(dp), q.shape[0]), "SetTrajectoryDetail")
            self._traj_detail = True
        idx = torch.empty(b.n, dtype=torch.int32, device=b.device) if want_index else None
        m = None if mask is None else mask.to(device=b.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(b.device):
            _lib.check(b.L.Cassie2dBatchEnvResetSampled(b.h, self.task, int(seed), int(first_global_env), int(draw),
                                                        None if m is None else m.data_ptr(), None if idx is None else idx.data_ptr(),
                                                        self._obs.data_ptr(), _stream_ptr()), "EnvResetSampled")
        return (self._obs, idx) if want_index else self._obs

    def step(self, action, n=10):
        """-> obs, reward, done.  With auto_reset a done env has already been reset when the call returns and its
        obs is the one env.reset() gives (so the next action is computed for the new episode); terminal_obs=True
        keeps the terminal observation instead.  An env whose state went non-finite reports done, reward 0, is
        reset (solver state included) and shows status 3 in stats()[:, 3]."""
        b = self.batch
        a = b._t(action, self.action_dim)
        with torch.cuda.device(b.device):
            _lib.check(b.L.Cassie2dBatchEnvStep(b.h, self.task, self.mode, a.data_ptr(), int(n), self.flags,
                                                self._obs.data_ptr(), self._rew.data_ptr(), self._done.data_ptr(),
                                                _stream_ptr()), "EnvStep")
        return self._obs, self._rew, self._done

    def step_host(self, action_host, obs_host, reward_host, done_host, n=10):
        """End-to-end variant: pinned host tensors in/out, copies + launch + sync inside the call."""
        b = self.batch
        _lib.check(b.L.Cassie2dBatchEnvStepHost(b.h, self.task, self.mode, action_host.data_ptr(), int(n), self.flags,
                                                obs_host.data_ptr(), reward_host.data_ptr(), done_host.data_ptr()), "EnvStepHost")

    def render(self):
        pass

    def terminate(self):
        self.batch.close()
