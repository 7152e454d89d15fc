"""long random-torque soak of the 3-D engine: states stay finite, quaternions unit, resets keep happening, capacity drops stay rare"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cassierl_b200.envs3d import Cassie3dBatch, TORQUE_HIGH_3D
n, launches = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 2000
b = Cassie3dBatch(n, precision=32)
g = torch.Generator(device="cuda").manual_seed(7)
hi = torch.tensor(TORQUE_HIGH_3D, dtype=torch.float32, device="cuda")
done2 = torch.zeros((), dtype=torch.int64, device="cuda")
rows_max = 0
t0 = time.time()
for k in range(launches):
    a = (torch.rand((n, 10), generator=g, device="cuda") * 2 - 1) * hi
    d = b.step(a, n=10, z_done=0.5, auto_reset=True)
    done2 += (d == 2).sum()
    if k % 200 == 0:
        rows_max = max(rows_max, int(b.stats()[:, 0].max().item()))
torch.cuda.synchronize()
q, v = b.state()
st = b.stats()
print("soak: %d envs x %d sim steps (%.3g env-steps) in %.1f s; finite %s; |quat|-1 max %.2e; non-finite events %d; auto-resets %d; "
      "envs that ever dropped a contact for capacity %d (contacts dropped %d); rows max seen %d; |qvel| max %.1f"
      % (n, launches * 10, n * launches * 10, time.time() - t0, bool(torch.isfinite(q).all() and torch.isfinite(v).all()),
         (q[:, 3:7].norm(dim=1) - 1).abs().max().item(), int(done2.item()), int(b.resets().sum().item()),
         int((st[:, 3] > 0).sum().item()), int(st[:, 3].sum().item()), rows_max, v.abs().max().item()))
