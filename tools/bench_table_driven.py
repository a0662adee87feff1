#!/usr/bin/env python
"""Panda objective through the TABLE-DRIVEN static kernel (k_objective_static): RobotPanda's model with one self pair
removed, so it is no compile-time collision model any more.  2,097,152 configurations."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_robotics_b200._lib import DeviceCost, DeviceModel  # noqa: E402
from torch_robotics_b200.workloads import panda_spheres3d_objective  # noqa: E402

desc, cost = panda_spheres3d_objective()
cost.pair_i, cost.pair_j, cost.pair_margin = cost.pair_i[:-1].copy(), cost.pair_j[:-1].copy(), cost.pair_margin[:-1].copy()
m = DeviceModel(desc, "cuda:0")
dc = DeviceCost(m, cost)
n = 1 << 21
lo, hi = desc.lo[desc.controlled], desc.hi[desc.controlled]
q = torch.as_tensor((lo + np.random.default_rng(1).random((n, 7), dtype=np.float32) * (hi - lo)).astype(np.float32), device="cuda:0")
tot, grad = torch.empty(n, device="cuda:0"), torch.empty(n, 7, device="cuda:0")
for _ in range(3):
    dc.objective(q, want_terms=False, out_total=tot, out_grad=grad)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    dc.objective(q, want_terms=False, out_total=tot, out_grad=grad)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(json.dumps({"kernel": "k_objective_static<Panda> (table-driven, 9 pairs)", "n": n, "ms": ms, "configs_per_s": n / ms * 1e3}))
