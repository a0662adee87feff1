"""Host-buffer entry point of the fused objective: pinned q in, pinned cost + grad_q out.

Thin wrapper over the C ABI call `trk_objective_host`, which cuts the batch into chunks that flow through three
CUDA streams inside the library - H2D copy, kernel, D2H copy - so the PCIe transfers in both directions overlap
each other and the compute.  This is the end-to-end path bench.py reports as `e2e` (host<->device copies inside
the timed region)."""
from __future__ import annotations

import torch

from ._lib import DeviceCost


class HostObjectivePipeline:
    def __init__(self, cost: DeviceCost, n_max: int, n_chunks: int = 8):
        self.cost, self.n_max, self.n_chunks = cost, int(n_max), max(1, int(n_chunks))

    def run(self, q_host: torch.Tensor, cost_host: torch.Tensor, grad_host: torch.Tensor | None = None) -> None:
        """q_host [n, D] pinned fp32 -> cost_host [n] (and grad_host [n, D]) pinned fp32.  Returns after the
        results are in host memory."""
        if q_host.shape[0] > self.n_max:
            raise RuntimeError(f"batch {q_host.shape[0]} exceeds the pipeline capacity {self.n_max}")
        for t, name in ((q_host, "q_host"), (cost_host, "cost_host"), (grad_host, "grad_host")):
            if t is not None and not t.is_pinned():
                raise RuntimeError(f"{name} must be pinned (page-locked) host memory")
        self.cost.objective_host(q_host, cost_host, grad_host, self.n_chunks)
