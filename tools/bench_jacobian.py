#!/usr/bin/env python
"""Time trk_link_jacobian (q [N,D] -> J [N,L,7,D]) and report it against the HBM roofline (output-bound:
L*7*D*4 B per configuration).  usage: python tools/bench_jacobian.py [panda|stretch|shadow_hand] [N]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_robotics_b200.torch_kinematics_tree.models import robots as R  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "panda"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    model = dict(panda=lambda: R.DifferentiableFrankaPanda(gripper=False, device="cuda:0"),
                 stretch=lambda: R.DifferentiableHabitatStretch(device="cuda:0"),
                 shadow_hand=lambda: R.DifferentiableShadowHand(device="cuda:0"))[name]()
    dm = model.device_model("cuda:0")
    q = torch.rand(n, dm.n_dofs, device="cuda:0") * 2 - 1
    for _ in range(3):
        J = dm.link_jacobian(q)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        J = dm.link_jacobian(q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_alg = n * (dm.n_links * 7 * dm.n_dofs * 4 + dm.n_dofs * 4)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6535.7) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6535.7
    print(json.dumps(dict(kernel="k_link_jacobian", robot=name, n=n, ms=ms, configs_per_s=n / ms * 1e3,
                          gbs=bytes_alg / ms / 1e6, hbm_frac=bytes_alg / ms / 1e6 / peak, bytes_per_config=bytes_alg // n)))


if __name__ == "__main__":
    main()
