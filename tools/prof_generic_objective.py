import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["TRK_B200_FORCE_GENERIC"] = "1"
from torch_robotics_b200._lib import DeviceCost, DeviceModel
from conftest import get_desc, load_golden, planning_cost_from_golden
d = get_desc("panda")
m = DeviceModel(d, "cuda:0")
cost = planning_cost_from_golden(load_golden("planning_panda_spheres3d"))
dc = DeviceCost(m, cost)
n = 1 << 21
lo, hi = d.lo[d.controlled], d.hi[d.controlled]
q = torch.as_tensor((lo + np.random.default_rng(1).random((n, 7), dtype=np.float32) * (hi - lo)).astype(np.float32), device="cuda:0")
tot, grad = torch.empty(n, device="cuda:0"), torch.empty(n, 7, device="cuda:0")
for _ in range(5):
    dc.objective(q, want_terms=False, out_total=tot, out_grad=grad)
torch.cuda.synchronize()
