#!/usr/bin/env python
"""Runtime-tree ("generic") kernel family: FK forward / backward on a URDF outside the compiled-in set
(iiwa7_allegro: 29 links, 23 DoF) and, with the compiled-in families switched off, Panda FK + fused objective."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from torch_robotics_b200._lib import DeviceCost, DeviceModel  # noqa: E402
from torch_robotics_b200.descriptor import compile_urdf  # noqa: E402

HBM = 6535.7


def timed(fn, inner=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    return float(np.median(ts))


def fk_lines(name, d, n, library=None):
    m = DeviceModel(d, "cuda:0", library=library)
    q = torch.as_tensor(np.random.default_rng(0).uniform(-1, 1, (n, d.n_dofs)).astype(np.float32), device="cuda:0")
    H = m.fk_forward(q)
    G = torch.randn_like(H)
    bf = n * (d.n_dofs * 4 + d.n_links * 64)
    bb = n * (2 * d.n_dofs * 4 + d.n_links * 64)
    msf, msb = timed(lambda: m.fk_forward(q)), timed(lambda: m.fk_backward(q, G))
    print(json.dumps({"robot": name, "family": m.kernel_family, "links": d.n_links, "dofs": d.n_dofs, "n": n,
                      "fk_forward": {"ms": msf, "GBps": bf / msf / 1e6, "hbm_frac": bf / msf / 1e6 / HBM},
                      "fk_backward": {"ms": msb, "GBps": bb / msb / 1e6, "hbm_frac": bb / msb / 1e6 / HBM}}))
    return m


def specialised_lines(d_iiwa):
    """The same robot through a library specialised at run time (specialize.py), and the UR10 with an ad-hoc collision model:
    default library (table-driven kernel) against the specialised one (headline kernel)."""
    import specialized_cases as sc
    from torch_robotics_b200 import specialize as sp
    from torch_robotics_b200.descriptor import build_cost_descriptor
    fk_lines("iiwa7_allegro (specialised)", d_iiwa, 1 << 19, library=sp.build_specialized_library(d_iiwa))
    d, cm = sc.ur10_collision()
    rng = np.random.default_rng(11)
    centers = rng.uniform(-0.7, 0.7, (10, 3)).astype(np.float32)
    cost = build_cost_descriptor(pt_link=cm["pt_link"], pt_radius=np.full(len(cm["pt_link"]), 0.08, np.float32), cutoff_obj=0.02,
                                 objects=[(centers, np.full(10, 0.12, np.float32), None, None)], ws_min=[-1.2, -1.2, -0.4], ws_max=[1.2, 1.2, 1.4],
                                 self_pairs=cm["pairs"], self_margin=[0.05], ee_link=cm["ee_body"], ee_target=np.eye(4, dtype=np.float32))
    n = 1 << 21
    q = torch.as_tensor(rng.uniform(-np.pi, np.pi, (n, d.n_dofs)).astype(np.float32), device="cuda:0")
    tot, grad = torch.empty(n, device="cuda:0"), torch.empty(n, d.n_dofs, device="cuda:0")
    for name, lib in (("default library", None), ("specialised", sp.build_specialized_library(d, cm))):
        dc = DeviceCost(DeviceModel(d, "cuda:0", library=lib), cost)
        ms = timed(lambda: dc.objective(q, want_terms=False, out_total=tot, out_grad=grad))
        print(json.dumps({"robot": f"ur10 + ad-hoc collision model ({name})", "n": n, "ms": ms, "configs_per_s": n / ms * 1e3}))


def main():
    urdf = os.path.join(ROOT, "torch_robotics_b200", "data", "urdf", "robots")
    d_iiwa = compile_urdf(os.path.join(urdf, "kuka_iiwa", "urdf", "iiwa7_allegro.urdf"), "iiwa7_allegro")
    fk_lines("iiwa7_allegro", d_iiwa, 1 << 19)
    specialised_lines(d_iiwa)
    os.environ["TRK_B200_FORCE_GENERIC"] = "1"
    from conftest import get_desc, load_golden, planning_cost_from_golden
    d = get_desc("panda")
    m = fk_lines("panda (forced generic)", d, 1 << 20)
    fk_lines("stretch (forced generic)", get_desc("stretch"), 1 << 19)
    cost = planning_cost_from_golden(load_golden("planning_panda_spheres3d"))
    dc = DeviceCost(m, cost)
    n = 1 << 21
    lo, hi = d.lo[d.controlled], d.hi[d.controlled]
    q = torch.as_tensor((lo + np.random.default_rng(1).random((n, 7), dtype=np.float32) * (hi - lo)).astype(np.float32), device="cuda:0")
    tot, grad = torch.empty(n, device="cuda:0"), torch.empty(n, 7, device="cuda:0")
    ms = timed(lambda: dc.objective(q, want_terms=False, out_total=tot, out_grad=grad))
    print(json.dumps({"robot": "panda (forced generic)", "kernel": "k_objective_generic", "n": n, "ms": ms, "configs_per_s": n / ms * 1e3,
                      "fp32_roofline_frac": n / ms * 1e3 * 2142 / 74.45e12}))


if __name__ == "__main__":
    main()
