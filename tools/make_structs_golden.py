"""Golden vectors for the FFI struct converters, generated with the REFERENCE's own
rllab/envs/cassie2d_structs.py (imported from /root/reference; ctypes + numpy only).  Run in the build
container; the test (tests/test_structs_golden.py) compares cassierl_b200.structs against them."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/rllab/envs")
import cassie2d_structs as ref  # noqa: E402

rng = np.random.default_rng(2018)
cv = ref.InterfaceStructConverter()
out = {}
s18 = rng.standard_normal((5, 18)); s26 = rng.standard_normal((5, 26)); a7 = rng.standard_normal((5, 7)); a6 = rng.standard_normal((5, 6))
out["s18"], out["s26"], out["a7"], out["a6"] = s18, s26, a7, a6
out["pos_invariant"] = np.array([cv.operational_state_array_to_pos_invariant_array(s) for s in s18])
gs = [cv.array_to_general_state(s) for s in s26]
out["general_roundtrip"] = np.array([cv.general_state_to_array(g) for g in gs])
out["general_bytes"] = np.array([np.frombuffer(bytes(g), np.float64) for g in gs])
out["osc_bytes"] = np.array([np.frombuffer(bytes(cv.array_to_operational_action(a)), np.float64) for a in a7])
out["torque_bytes"] = np.array([np.frombuffer(bytes(cv.array_to_torque_action(a)), np.float64) for a in a6])
out["pd_bytes"] = np.array([np.frombuffer(bytes(cv.array_to_pd_action(a)), np.float64) for a in a6])
xs = []
for s in s18:
    x = ref.StateOperationalSpace()
    for i in range(3):
        x.body_x[i], x.body_xd[i], x.left_x[i], x.left_xd[i], x.right_x[i], x.right_xd[i] = s[i], s[3 + i], s[6 + i], s[9 + i], s[12 + i], s[15 + i]
    xs.append(cv.operational_state_to_array(x))
out["op_to_array"] = np.array(xs)
out["sizes"] = np.array([ctypes.sizeof(getattr(ref, n)) for n in
                         ("ControllerTorque", "ControllerForce", "ControllerOsc", "ControllerPd", "StateGeneral", "StateOperationalSpace")])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "structs_reference.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
