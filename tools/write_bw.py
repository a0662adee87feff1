#!/usr/bin/env python
"""Pure-write and pure-read HBM bandwidth next to the copy bandwidth MEASURED_PEAKS.json quotes: FK forward is ~96 %
writes, FK backward ~96 % reads, so these are the ceilings that actually apply to them."""
import json

import torch

n = 768 * 1024 * 1024 // 4
x = torch.empty(n, device="cuda", dtype=torch.float32)
y = torch.empty(n, device="cuda", dtype=torch.float32)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ms_w = timed(lambda: x.zero_())
ms_c = timed(lambda: y.copy_(x))
ms_r = timed(lambda: x.sum())
print(json.dumps({"bytes": n * 4, "write_GBps": n * 4 / ms_w / 1e6, "copy_GBps": 2 * n * 4 / ms_c / 1e6, "read_GBps(sum)": n * 4 / ms_r / 1e6}))
