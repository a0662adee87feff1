#!/usr/bin/env python
"""Launch Panda FK forward / backward at 1,048,576 configurations a few times (target for ncu captures)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import get_desc  # noqa: E402
from torch_robotics_b200._lib import DeviceModel  # noqa: E402

d = get_desc(sys.argv[1] if len(sys.argv) > 1 else "panda")
m = DeviceModel(d, "cuda:0")
n = 1 << 20
q = torch.as_tensor(np.random.default_rng(0).uniform(-1, 1, (n, d.n_dofs)).astype(np.float32), device="cuda:0")
H = m.fk_forward(q)
G = torch.randn_like(H)
for _ in range(4):
    m.fk_forward(q)
    m.fk_backward(q, G)
torch.cuda.synchronize()
