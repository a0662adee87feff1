"""Host-buffer entry point of the fused objective: pinned q in, pinned cost + grad_q out.

The batch is cut into chunks that flow through three CUDA streams - H2D copy, kernel, D2H copy - so
the PCIe transfers in both directions overlap each other and the compute.  This is the end-to-end path
bench.py reports as `e2e` (host<->device copies inside the timed region)."""
from __future__ import annotations

import torch

from ._lib import DeviceCost


class HostObjectivePipeline:
    def __init__(self, cost: DeviceCost, n_max: int, n_chunks: int = 8):
        self.cost, self.n_max = cost, int(n_max)
        dev = cost.model.device
        D = cost.model.n_dofs
        self.dev = dev
        self.q_dev = torch.empty((n_max, D), dtype=torch.float32, device=dev)
        self.total_dev = torch.empty(n_max, dtype=torch.float32, device=dev)
        self.grad_dev = torch.empty((n_max, D), dtype=torch.float32, device=dev)
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.n_chunks = max(1, int(n_chunks))

    def run(self, q_host: torch.Tensor, cost_host: torch.Tensor, grad_host: torch.Tensor | None = None) -> None:
        """q_host [n, D] pinned fp32 -> cost_host [n] (and grad_host [n, D]) pinned fp32.  Returns after the
        results are in host memory."""
        n = q_host.shape[0]
        if n > self.n_max:
            raise RuntimeError(f"batch {n} exceeds the pipeline capacity {self.n_max}")
        for t, name in ((q_host, "q_host"), (cost_host, "cost_host"), (grad_host, "grad_host")):
            if t is not None and (t.is_cuda or not t.is_pinned() or t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError(f"{name} must be a contiguous pinned float32 host tensor")
        caller = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(caller)
        step = -(-n // self.n_chunks)
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            with torch.cuda.stream(self.s_in):
                self.q_dev[lo:hi].copy_(q_host[lo:hi], non_blocking=True)
            self.s_k.wait_stream(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.cost.objective(self.q_dev[lo:hi], want_grad=grad_host is not None, want_terms=False,
                                    out_total=self.total_dev[lo:hi],
                                    out_grad=self.grad_dev[lo:hi] if grad_host is not None else None)
            self.s_out.wait_stream(self.s_k)
            with torch.cuda.stream(self.s_out):
                cost_host[lo:hi].copy_(self.total_dev[lo:hi], non_blocking=True)
                if grad_host is not None:
                    grad_host[lo:hi].copy_(self.grad_dev[lo:hi], non_blocking=True)
        self.s_out.synchronize()
