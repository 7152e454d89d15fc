"""Per-call latency of the legacy one-env ABI: StepTorque / StepOsc / GetGeneralState."""
import ctypes as ct
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200 import lib, structs as S

L = lib.load()
h = L.Cassie2dInit()
t = S.ControllerTorque(); o = S.ControllerOsc(); q = S.StateGeneral()
for name, fn, arg in (("StepTorque", L.StepTorque, t), ("StepOsc", L.StepOsc, o), ("GetGeneralState", L.GetGeneralState, q)):
    for _ in range(50):
        fn(h, ct.byref(arg))
    t0 = time.perf_counter()
    for _ in range(1000):
        fn(h, ct.byref(arg))
    print("%-16s %.1f us per call (fp64, batch of one)" % (name, (time.perf_counter() - t0) * 1e3))
