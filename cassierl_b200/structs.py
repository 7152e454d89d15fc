"""ctypes mirror of the FFI structs (CassieRL/cassierl src/Cassie2d/RobotInterface.h:14-50 ==
rllab/envs/cassie2d_structs.py:5-51) and the array<->struct converters (:52-122), with the same
names so that code written against the reference's `cassie2d_structs` runs unchanged."""
import ctypes

import numpy as np


class ControllerTorque(ctypes.Structure):
    _fields_ = [("torques", ctypes.c_double * 6)]


class ControllerForce(ctypes.Structure):
    _fields_ = [("left_force", ctypes.c_double * 3), ("right_force", ctypes.c_double * 3)]


class ControllerOsc(ctypes.Structure):
    _fields_ = [("body_xdd", ctypes.c_double * 2), ("left_xdd", ctypes.c_double * 2),
                ("right_xdd", ctypes.c_double * 2), ("pitch_add", ctypes.c_double)]


class ControllerPd(ctypes.Structure):
    _fields_ = [("angles", ctypes.c_double * 6)]


class StateGeneral(ctypes.Structure):
    _fields_ = [("base_pos", ctypes.c_double * 3), ("base_vel", ctypes.c_double * 3),
                ("left_pos", ctypes.c_double * 5), ("left_vel", ctypes.c_double * 5),
                ("right_pos", ctypes.c_double * 5), ("right_vel", ctypes.c_double * 5)]


class StateOperationalSpace(ctypes.Structure):
    _fields_ = [("body_x", ctypes.c_double * 3), ("body_xd", ctypes.c_double * 3),
                ("left_x", ctypes.c_double * 3), ("left_xd", ctypes.c_double * 3),
                ("right_x", ctypes.c_double * 3), ("right_xd", ctypes.c_double * 3)]


def _flat(struct):
    return np.frombuffer(struct, dtype=np.float64).copy()


class InterfaceStructConverter:
    """Same methods as the reference's converter; all structs are flat double arrays."""

    def operational_state_to_array(self, state):
        return _flat(state)                      # body_x body_xd left_x left_xd right_x right_xd

    def operational_state_array_to_pos_invariant_array(self, state_array):
        s = np.zeros((26,), dtype=np.float64)
        s[:17] = state_array[1:18]
        s[5] -= state_array[0]
        s[11] -= state_array[0]
        return s

    def general_state_to_array(self, state):
        return _flat(state)

    def array_to_general_state(self, s):
        return StateGeneral.from_buffer_copy(np.ascontiguousarray(s[:26], np.float64).tobytes())

    def array_to_operational_action(self, action):
        return ControllerOsc.from_buffer_copy(np.ascontiguousarray(action[:7], np.float64).tobytes())

    def array_to_torque_action(self, action):
        return ControllerTorque.from_buffer_copy(np.ascontiguousarray(action[:6], np.float64).tobytes())

    def array_to_pd_action(self, action):
        return ControllerPd.from_buffer_copy(np.ascontiguousarray(action[:6], np.float64).tobytes())
