#!/usr/bin/env python
"""Runtime-tree ("generic") kernel family on a URDF outside the compiled-in set (iiwa7_allegro: 29 links, 23 DoF)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_robotics_b200._lib import DeviceModel
from torch_robotics_b200.descriptor import compile_urdf
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = compile_urdf(os.path.join(root, "torch_robotics_b200", "data", "urdf", "robots", "kuka_iiwa", "urdf", "iiwa7_allegro.urdf"), "iiwa7_allegro")
m = DeviceModel(d, "cuda:0")
n = 1 << 19
q = torch.as_tensor(np.random.default_rng(0).uniform(-1, 1, (n, d.n_dofs)).astype(np.float32), device="cuda:0")
H = m.fk_forward(q); G = torch.randn_like(H)
def timed(fn, inner=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner): fn()
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / inner)
    return float(np.median(ts))
bf = n * (d.n_dofs * 4 + d.n_links * 64); bb = n * (2 * d.n_dofs * 4 + d.n_links * 64)
msf, msb = timed(lambda: m.fk_forward(q)), timed(lambda: m.fk_backward(q, G))
print(json.dumps({"family": m.kernel_family, "links": d.n_links, "dofs": d.n_dofs, "n": n,
                  "fk_forward": {"ms": msf, "GBps": bf / msf / 1e6, "hbm_frac": bf / msf / 1e6 / 6535.7},
                  "fk_backward": {"ms": msb, "GBps": bb / msb / 1e6, "hbm_frac": bb / msb / 1e6 / 6535.7}}))
