#!/usr/bin/env python
"""Kernel-only timing of the fused Panda objective (development aid): prints ms and configs/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch_robotics_b200._lib import DeviceCost, DeviceModel, LIB_PATH
from torch_robotics_b200.workloads import panda_spheres3d_objective
desc, cost = panda_spheres3d_objective()
m = DeviceModel(desc, "cuda:0"); dc = DeviceCost(m, cost)
n = 4194304
rng = np.random.default_rng(0)
lo, hi = desc.lo[desc.controlled], desc.hi[desc.controlled]
q = torch.as_tensor((lo + rng.random((n, 7), dtype=np.float32) * (hi - lo)).astype(np.float32), device="cuda:0")
tot = torch.empty(n, device="cuda:0"); g = torch.empty_like(q)
for _ in range(5): dc.objective(q, want_terms=False, out_total=tot, out_grad=g)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dc.objective(q, want_terms=False, out_total=tot, out_grad=g); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
print(json.dumps({"lib": os.path.basename(LIB_PATH), "family": m.kernel_family, "ms": ms, "cfg_per_s": n / ms * 1e3, "frac_fp32": n / ms * 1e3 * 2142 / 74.45e12}))
