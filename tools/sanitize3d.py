"""small 3-D run for compute-sanitizer (memcheck / racecheck): ragged batch, contacts, limits, auto-reset, every lane width"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200.envs3d import Cassie3dBatch, TORQUE_HIGH_3D
n = 37
for lanes in (32, 16, 8):
    b = Cassie3dBatch(n, precision=32, lanes=lanes)
    q, v = b.state()
    q[:, 2] = 0.94                      # toes on the floor
    q[::3, 2] = 0.45; q[::3, 3:7] = torch.tensor([0.8253, 0.5646, 0.0, 0.0])    # some robots on their side: many contacts
    b.set_state(q, v)
    g = torch.Generator(device="cuda").manual_seed(lanes)
    hi = torch.tensor(TORQUE_HIGH_3D, dtype=torch.float32, device="cuda")
    for k in range(6):
        a = (torch.rand((n, 10), generator=g, device="cuda") * 2 - 1) * hi
        b.step(a, n=5, z_done=0.5, auto_reset=True)
    torch.cuda.synchronize()
    st = b.stats().cpu().numpy()
    print("lanes", lanes, "rows max", st[:, 0].max(), "contacts max", st[:, 1].max(), "resets", int(b.resets().sum().item()), "finite", bool(torch.isfinite(b.state()[0]).all()))
    b.close()
