"""The C-ABI library loads and exports every symbol include/trk_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from torch_robotics_b200 import _lib
    header = open(os.path.join(ROOT, "include", "trk_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(trk_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.lib().trk_abi_version() == 1


def test_invalid_arguments_return_status_not_crash():
    from torch_robotics_b200 import _lib
    L = _lib.lib()
    out = ctypes.c_void_p()
    assert L.trk_model_create(None, 0, ctypes.byref(out)) == -1
    assert b"desc" in L.trk_last_error()
    md = _lib.ModelDesc()
    md.n_links = 1000
    assert L.trk_model_create(ctypes.byref(md), 0, ctypes.byref(out)) == -2
    assert L.trk_fk_forward(None, None, 0, None, 0, None, None) == -1
    assert L.trk_link_jacobian(None, None, 0, None, None) == -1
    assert b"model" in L.trk_last_error()
    assert L.trk_cost_create(None, None, ctypes.byref(out)) == -1
    assert L.trk_objective_host(None, None, None, 0, 8, None, None) == -1
