#!/usr/bin/env python
"""Ceiling for the host-buffer objective call: H2D of q (117 MB) and D2H of cost + grad (134 MB) running at the same
time on two streams, no kernel in between."""
import json
import time

import torch

n = 4194304
hq = torch.empty(n, 7, pin_memory=True)
hg = torch.empty(n, 8, pin_memory=True)
dq = torch.empty(n, 7, device="cuda")
dg = torch.empty(n, 8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d=True, d2h=True):
    if h2d:
        with torch.cuda.stream(s1):
            dq.copy_(hq, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2):
            hg.copy_(dg, non_blocking=True)
    torch.cuda.synchronize()


out = {}
for name, kw in (("h2d_only", dict(d2h=False)), ("d2h_only", dict(h2d=False)), ("duplex", {})):
    for _ in range(3):
        run(**kw)
    t0 = time.perf_counter()
    for _ in range(10):
        run(**kw)
    dt = (time.perf_counter() - t0) / 10
    nbytes = (hq.numel() * 4 if kw.get("h2d", True) else 0) + (hg.numel() * 4 if kw.get("d2h", True) else 0)
    out[name] = {"ms": dt * 1e3, "GBps": nbytes / dt / 1e9}
out["duplex_configs_per_s_ceiling"] = n / (out["duplex"]["ms"] * 1e-3)
print(json.dumps(out))
