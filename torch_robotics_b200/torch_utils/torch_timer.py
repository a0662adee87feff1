"""TimerCUDA context manager (reference: torch_utils/torch_timer.py:7-108): wall clock bracketed by
torch.cuda.synchronize(), or CUDA events on the current stream when use_cuda_events=True."""
from time import perf_counter

import torch


class TimerCUDA:
    def __init__(self, timer=perf_counter, output=None, fmt="took {:.3f} seconds", prefix="", use_cuda_events=False):
        self.timer, self.output, self.fmt, self.prefix = timer, output, fmt, prefix
        self.sync_cuda = torch.cuda.is_available()
        if use_cuda_events and not self.sync_cuda:
            raise AssertionError("CUDA must be available when using CUDA events")
        self.use_cuda_events = use_cuda_events
        self._t0 = self._t1 = None
        self._e0 = self._e1 = None

    def _now(self):
        if self.sync_cuda:
            torch.cuda.synchronize()
        return self.timer()

    def __enter__(self):
        if self.use_cuda_events:
            self._e0 = torch.cuda.Event(enable_timing=True)
            self._e1 = torch.cuda.Event(enable_timing=True)
            self._e0.record()
        else:
            self._t0 = self._now()
        return self

    def __exit__(self, *exc):
        if self.use_cuda_events:
            self._e1.record()
            torch.cuda.synchronize()
            self._t1 = self._e0.elapsed_time(self._e1) / 1000.0
        else:
            self._t1 = self._now()
        if self.prefix and self.output is None:
            self.output = True
        if self.output:
            msg = " ".join([self.prefix, self.fmt.format(self.elapsed)])
            self.output(msg) if callable(self.output) else print(msg)

    @property
    def elapsed(self):
        if self.use_cuda_events:
            if self._t1 is not None:
                return self._t1
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            torch.cuda.synchronize()
            return self._e0.elapsed_time(e) / 1000.0
        return (self._now() if self._t1 is None else self._t1) - self._t0

    def __str__(self):
        return f"{self.elapsed:.3f}"
