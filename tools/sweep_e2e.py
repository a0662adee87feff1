#!/usr/bin/env python
"""e2e (pinned host buffers -> cost + grad_q in pinned host buffers) vs number of pipeline chunks."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch_robotics_b200._lib import DeviceCost, DeviceModel
from torch_robotics_b200.host_pipeline import HostObjectivePipeline
from torch_robotics_b200.workloads import panda_spheres3d_objective
desc, cost = panda_spheres3d_objective()
m = DeviceModel(desc, "cuda:0"); dc = DeviceCost(m, cost)
n = 4194304
q_host = (torch.rand(n, 7) * 2 - 1).pin_memory()
c_host = torch.empty(n).pin_memory(); g_host = torch.empty(n, 7).pin_memory()
for chunks in (1, 2, 4, 8, 16, 32, 64):
    pipe = HostObjectivePipeline(dc, n, n_chunks=chunks)
    for _ in range(2): pipe.run(q_host, c_host, g_host)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): pipe.run(q_host, c_host, g_host)
    dt = (time.perf_counter() - t0) / 10
    print(json.dumps({"chunks": chunks, "ms": dt * 1e3, "cfg_per_s": n / dt, "GBps_total": n * 60 / dt / 1e9}))
# copy-only ceilings
d = torch.empty(n, 7, device="cuda:0"); torch.cuda.synchronize()
for name, fn in (("h2d_only", lambda: d.copy_(q_host, non_blocking=True)), ("d2h_only", lambda: g_host.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(json.dumps({"copy": name, "GBps": n * 28 / dt / 1e9}))
