#!/usr/bin/env python
"""BASELINE config 5: Habitat Stretch (44 links, 14 DoF, 6 branch nodes) FK and fused objective (15 collision
points x 10 spheres, ws bounds, EE) with gradient.  Prints one JSON line per kernel (rank 0 only, 1 GPU)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_robotics_b200._lib import DeviceCost, DeviceModel
from torch_robotics_b200.descriptor import build_cost_descriptor
from torch_robotics_b200.environments.env_spheres_3d import SPHERE_CENTERS, SPHERE_RADIUS
from torch_robotics_b200.torch_kinematics_tree.models.robots import DifferentiableHabitatStretch
from torch_robotics_b200.workloads import ik_example_target

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
tree = DifferentiableHabitatStretch(device="cuda:0")
d = tree._desc
m = tree.device_model("cuda:0")
names = d.link_names
pts = list(d.controlled) + [names.index("link_straight_gripper")]
cost = build_cost_descriptor(pt_link=pts, pt_radius=[0.05] * len(pts), cutoff_obj=0.01,
                             objects=[(np.array(SPHERE_CENTERS, np.float32), np.full(10, SPHERE_RADIUS, np.float32), None, None)],
                             ws_min=[-1, -1, -1], ws_max=[1, 1, 1], ee_link=names.index("link_straight_gripper"), ee_target=ik_example_target())
dc = DeviceCost(m, cost)
rng = np.random.default_rng(0)
lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float32)
hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float32)
q = torch.as_tensor((lo + rng.random((n, d.n_dofs), dtype=np.float32) * (hi - lo)).astype(np.float32), device="cuda:0")

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
tot, g = torch.empty(n, device="cuda:0"), torch.empty_like(q)
ms = timed(lambda: dc.objective(q, want_terms=False, out_total=tot, out_grad=g))
print(json.dumps({"kernel": "stretch_objective_fwd_bwd", "family": m.kernel_family, "n": n, "ms": ms, "configs_per_s": n / ms * 1e3,
                  "frac_fp32_roofline(5880 flop)": n / ms * 1e3 * 5880 / 74.45e12}))
n2 = min(n, 1 << 20)
H = m.fk_forward(q[:n2])
msf = timed(lambda: m.fk_forward(q[:n2]))
print(json.dumps({"kernel": "stretch_fk_forward", "n": n2, "ms": msf, "configs_per_s": n2 / msf * 1e3, "GBps": n2 * 2872 / msf / 1e6,
                  "hbm_frac": n2 * 2872 / msf / 1e6 / 6535.7}))
