"""cassierl_b200.structs against golden vectors produced by the reference's own
rllab/envs/cassie2d_structs.py (tools/make_structs_golden.py): struct sizes, field order, converters."""
import ctypes
import os

import numpy as np

from conftest import ROOT
from cassierl_b200 import structs as S


def test_struct_converters_match_reference_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "structs_reference.npz"))
    cv = S.InterfaceStructConverter()
    names = ("ControllerTorque", "ControllerForce", "ControllerOsc", "ControllerPd", "StateGeneral", "StateOperationalSpace")
    assert [ctypes.sizeof(getattr(S, n)) for n in names] == list(g["sizes"])
    for k in range(5):
        assert np.array_equal(cv.operational_state_array_to_pos_invariant_array(g["s18"][k]), g["pos_invariant"][k])
        st = cv.array_to_general_state(g["s26"][k])
        assert np.array_equal(np.frombuffer(bytes(st), np.float64), g["general_bytes"][k])
        assert np.array_equal(cv.general_state_to_array(st), g["general_roundtrip"][k])
        assert np.array_equal(np.frombuffer(bytes(cv.array_to_operational_action(g["a7"][k])), np.float64), g["osc_bytes"][k])
        assert np.array_equal(np.frombuffer(bytes(cv.array_to_torque_action(g["a6"][k])), np.float64), g["torque_bytes"][k])
        assert np.array_equal(np.frombuffer(bytes(cv.array_to_pd_action(g["a6"][k])), np.float64), g["pd_bytes"][k])
        x = S.StateOperationalSpace()
        s = g["s18"][k]
        for i in range(3):
            x.body_x[i], x.body_xd[i], x.left_x[i], x.left_xd[i], x.right_x[i], x.right_xd[i] = s[i], s[3 + i], s[6 + i], s[9 + i], s[12 + i], s[15 + i]
        assert np.array_equal(cv.operational_state_to_array(x), g["op_to_array"][k])
