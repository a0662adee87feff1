"""Drop-in alias: scripts written against the reference (`import torch_robotics....`) run on this package unchanged.

    import torch_robotics_b200.compat as compat
    compat.install_torch_robotics_alias()            # `torch_robotics.X.Y` now resolves to `torch_robotics_b200.X.Y`

The reference's example scripts use device="cpu" (examples/forward_kinematics.py:15).  This implementation computes on
CUDA only, so by default a CPU tensor raises.  `accept_cpu_tensors(True)` (or TRK_B200_ACCEPT_CPU_TENSORS=1) is an explicit
opt-in compatibility mode for such scripts: trees built with device="cpu" live on cuda:0, CPU inputs are copied to the GPU,
evaluated by the same CUDA kernels, and the results are copied back (autograd follows the copies).  There is still no
CPU implementation: without a GPU and the CUDA library the calls fail as before.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import os
import sys

_ACCEPT_CPU = os.environ.get("TRK_B200_ACCEPT_CPU_TENSORS", "") not in ("", "0")


def accept_cpu_tensors(flag: bool = True) -> None:
    global _ACCEPT_CPU
    _ACCEPT_CPU = bool(flag)


def cpu_tensors_accepted() -> bool:
    return _ACCEPT_CPU


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    SRC, DST = "torch_robotics", "torch_robotics_b200"

    def find_spec(self, fullname, path=None, target=None):
        if fullname == self.SRC or fullname.startswith(self.SRC + "."):
            real = self.DST + fullname[len(self.SRC):]
            try:
                if importlib.util.find_spec(real) is None:
                    return None
            except ModuleNotFoundError:
                return None
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        real = self.DST + spec.name[len(self.SRC):]
        return importlib.import_module(real)        # the very same module object under both names

    def exec_module(self, module):
        pass


def install_torch_robotics_alias() -> None:
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder())
