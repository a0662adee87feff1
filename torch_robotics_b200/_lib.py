"""ctypes binding of the C ABI in include/trk_b200.h (csrc/libtrk_b200.so).

There is deliberately no fallback: if the CUDA library is missing or a tensor is not a contiguous
fp32 CUDA tensor the call raises.  Build the library with `python -c "import __graft_entry__ as g;
g.build()"` (or `make -C torch_robotics_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRK_B200_LIB") or os.path.join(_HERE, "csrc", "libtrk_b200.so")   # env override: development builds

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("n_dofs", C.c_int32),
                ("parent", c_int32_p), ("kind", c_int32_p), ("axis_k", c_int32_p), ("axis_s", c_float_p),
                ("axis", c_float_p), ("q_index", c_int32_p), ("R_fixed", c_float_p), ("trans", c_float_p),
                ("lo", c_float_p), ("hi", c_float_p), ("has_lim", c_int32_p), ("topo", c_int32_p),
                ("base_pose", c_float_p)]


class CostDesc(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("pt_link", c_int32_p), ("pt_margin_obj", c_float_p), ("pt_margin_ws", c_float_p),
                ("n_objects", C.c_int32), ("obj_start", c_int32_p), ("obj_identity", c_int32_p), ("obj_R", c_float_p),
                ("obj_pos", c_float_p), ("n_spheres", C.c_int32), ("sph_center", c_float_p), ("sph_radius", c_float_p),
                ("n_boxes", C.c_int32), ("obj_box_start", c_int32_p), ("box_center", c_float_p), ("box_half", c_float_p),
                ("box_radius", c_float_p), ("box_rounded", c_int32_p),
                ("grid_obj", C.c_int32), ("grid_n", C.c_int32 * 3), ("grid_lo", C.c_float * 3), ("grid_extent", C.c_float * 3),
                ("grid_data", C.c_void_p),
                ("ws_enabled", C.c_int32), ("ws_min", C.c_float * 3), ("ws_max", C.c_float * 3),
                ("n_pairs", C.c_int32), ("pair_i", c_int32_p), ("pair_j", c_int32_p), ("pair_margin", c_float_p),
                ("ee_enabled", C.c_int32), ("ee_link", C.c_int32), ("ee_w_pos", C.c_float), ("ee_w_rot", C.c_float),
                ("ee_square", C.c_int32), ("ee_target", C.c_float * 16),
                ("clamp_self", C.c_int32), ("clamp_obj", C.c_int32), ("clamp_ws", C.c_int32), ("weights", C.c_float * 4)]


class IKParamsC(C.Structure):
    _fields_ = [("ee_link", C.c_int32), ("n_iters", C.c_int32), ("n_updates", C.c_int32),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("w_se3", C.c_float), ("w_joint_limits", C.c_float), ("se3_eps", C.c_float),
                ("lower", c_float_p), ("upper", c_float_p), ("H_target", c_float_p)]


_lib = None
_libs = {}                      # path -> configured CDLL (the default library and run-time specialised ones: specialize.py)
_lib_lock = threading.Lock()

EXPORTS = [
    "trk_abi_version", "trk_last_error", "trk_model_create", "trk_model_destroy", "trk_model_kernel_family",
    "trk_fk_forward", "trk_fk_backward", "trk_cost_create", "trk_cost_destroy", "trk_cost_set_ee_target",
    "trk_cost_set_weights", "trk_objective", "trk_objective_host", "trk_cost_stage", "trk_selftest_sqrt", "trk_ik_adam",
    "trk_link_jacobian", "trk_geometric_jacobian", "trk_objective_sums", "trk_ik_adam_best",
]


def load_library(path: str) -> C.CDLL:
    """dlopen a build of the C ABI (include/trk_b200.h) and declare its signatures; one handle per path."""
    path = os.path.abspath(path)
    L = _libs.get(path)
    if L is not None:
        return L
    with _lib_lock:
        L = _libs.get(path)
        if L is not None:
            return L
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: the sm_100a CUDA library has not been built. "
                "Run `python -c \"import __graft_entry__ as g; g.build()\"` in the repository root. "
                "There is no CPU or PyTorch fallback for this path.")
        L = C.CDLL(path)
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        L.trk_abi_version.restype = C.c_int
        L.trk_last_error.restype = C.c_char_p
        L.trk_model_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(vp)]
        L.trk_model_destroy.argtypes = [vp]
        L.trk_model_destroy.restype = None
        L.trk_model_kernel_family.argtypes = [vp]
        L.trk_model_kernel_family.restype = C.c_char_p
        L.trk_fk_forward.argtypes = [vp, vp, i64, c_int32_p, i32, vp, vp]
        L.trk_fk_backward.argtypes = [vp, vp, i64, c_int32_p, i32, vp, vp, vp]
        L.trk_cost_create.argtypes = [vp, C.POINTER(CostDesc), C.POINTER(vp)]
        L.trk_cost_destroy.argtypes = [vp]
        L.trk_cost_destroy.restype = None
        L.trk_cost_set_ee_target.argtypes = [vp, c_float_p, vp]
        L.trk_cost_set_weights.argtypes = [vp, c_float_p, vp]
        L.trk_objective.argtypes = [vp, vp, vp, i64, vp, f32, vp, vp, vp, vp, vp, vp]
        L.trk_objective_sums.argtypes = [vp, vp, vp, i64, vp, f32, vp, vp, vp, vp]
        L.trk_cost_stage.argtypes = [vp, vp, vp, vp, i64, vp, f32, vp, vp, vp, vp, vp, vp, vp]
        L.trk_selftest_sqrt.argtypes = [vp, i64, vp, vp, vp]
        L.trk_link_jacobian.argtypes = [vp, vp, i64, vp, vp]
        L.trk_geometric_jacobian.argtypes = [vp, vp, i64, i32, C.c_uint32, vp, vp, vp, vp]
        L.trk_objective_host.argtypes = [vp, vp, vp, i64, i32, vp, vp]
        L.trk_ik_adam.argtypes = [vp, C.POINTER(IKParamsC), vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]
        L.trk_ik_adam_best.argtypes = [vp, C.POINTER(IKParamsC), vp, vp, vp, i64, vp, vp, vp, vp, vp, i32, vp]
        for name in EXPORTS:
            getattr(L, name)              # every declared entry point must be exported
        if L.trk_abi_version() != 1:
            raise ImportError(f"{path}: ABI version mismatch")
        _libs[path] = L
        return L


def lib() -> C.CDLL:
    """The default library (all shipped robots compiled in)."""
    global _lib
    if _lib is None:
        _lib = load_library(LIB_PATH)
    return _lib


def check(rc: int, what: str, L: Optional[C.CDLL] = None) -> None:
    if rc != 0:
        msg = (L or lib()).trk_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def _fp(a: np.ndarray):
    return a.ctypes.data_as(c_float_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(c_int32_p)


def _c32(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype))


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not torch.is_tensor(t):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} lives on {t.device}: the B200-native path only accepts CUDA tensors "
                           "(there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class DeviceModel:
    """Device-resident kinematic tree (trk_model) for one (descriptor, device)."""

    def __init__(self, desc, device, library: Optional[str] = None):
        """library: path of a specialised build of the C ABI for this robot (specialize.build_specialized_library);
        default: the package's library with the shipped robots compiled in."""
        self.desc = desc
        self._L = load_library(library) if library else lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"cannot create a device model on '{device}': CUDA only")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        torch.cuda.init()
        self._keep = dict(
            parent=_c32(desc.parent, np.int32), kind=_c32(desc.kind, np.int32), axis_k=_c32(desc.axis_k, np.int32),
            axis_s=_c32(desc.axis_s, np.float32), axis=_c32(desc.axis, np.float32), q_index=_c32(desc.q_index, np.int32),
            R_fixed=_c32(desc.R_fixed, np.float32), trans=_c32(desc.trans, np.float32), lo=_c32(desc.lo, np.float32),
            hi=_c32(desc.hi, np.float32), has_lim=_c32(desc.has_lim, np.int32), topo=_c32(desc.topo, np.int32),
            base=_c32(np.concatenate([np.asarray(desc.base_R, np.float32).reshape(9), np.asarray(desc.base_t, np.float32).reshape(3)]), np.float32))
        k = self._keep
        md = ModelDesc(desc.n_links, desc.n_dofs, _ip(k["parent"]), _ip(k["kind"]), _ip(k["axis_k"]), _fp(k["axis_s"]),
                       _fp(k["axis"]), _ip(k["q_index"]), _fp(k["R_fixed"]), _fp(k["trans"]), _fp(k["lo"]), _fp(k["hi"]),
                       _ip(k["has_lim"]), _ip(k["topo"]), _fp(k["base"]))
        h = C.c_void_p()
        check(self._L.trk_model_create(C.byref(md), idx, C.byref(h)), "trk_model_create", self._L)
        self.handle = h
        self.n_links, self.n_dofs = desc.n_links, desc.n_dofs
        self._ik_schedules = {}

    @property
    def kernel_family(self) -> str:
        return self._L.trk_model_kernel_family(self.handle).decode()

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and getattr(self, "_L", None) is not None:
            try:
                self._L.trk_model_destroy(h)
            except Exception:
                pass
            self.handle = None

    # -- raw calls (no autograd) ---------------------------------------------------------------
    def _sel(self, link_sel: Optional[Sequence[int]]):
        if link_sel is None:
            return None, self.n_links, None
        arr = np.ascontiguousarray(np.asarray(link_sel, np.int32))
        return _ip(arr), int(arr.shape[0]), arr

    def fk_forward(self, q: torch.Tensor, link_sel=None) -> torch.Tensor:
        q = require_cuda_f32(q, "q")
        if q.ndim != 2 or q.shape[1] != self.n_dofs:
            raise RuntimeError(f"q must be [N, {self.n_dofs}], got {tuple(q.shape)}")
        if q.device != self.device:
            raise RuntimeError(f"q is on {q.device} but the model lives on {self.device}")
        p, n_sel, keep = self._sel(link_sel)
        H = torch.empty((q.shape[0], n_sel, 4, 4), dtype=torch.float32, device=q.device)
        check(self._L.trk_fk_forward(self.handle, q.data_ptr(), q.shape[0], p, n_sel, H.data_ptr(), stream_ptr(q.device)),
              "trk_fk_forward", self._L)
        return H

    def link_jacobian(self, q: torch.Tensor) -> torch.Tensor:
        """q [N, D] -> J [N, L, 7, D] = d (position, quaternion wxyz) of every body / d q (trk_link_jacobian)."""
        q = require_cuda_f32(q, "q")
        if q.ndim != 2 or q.shape[1] != self.n_dofs:
            raise RuntimeError(f"q must be [N, {self.n_dofs}], got {tuple(q.shape)}")
        if q.device != self.device:
            raise RuntimeError(f"q is on {q.device} but the model lives on {self.device}")
        J = torch.empty((q.shape[0], self.n_links, 7, self.n_dofs), dtype=torch.float32, device=q.device)
        check(self._L.trk_link_jacobian(self.handle, q.data_ptr(), q.shape[0], J.data_ptr(), stream_ptr(q.device)), "trk_link_jacobian", self._L)
        return J

    def geometric_jacobian(self, q: torch.Tensor, link: int, col_mask: int):
        """q [N, D] -> (pose [N, 12] = R row-major | p of body `link`, lin [N, 3, D], ang [N, 3, D]) for the q columns in
        `col_mask` (trk_geometric_jacobian)."""
        q = require_cuda_f32(q, "q")
        if q.ndim != 2 or q.shape[1] != self.n_dofs:
            raise RuntimeError(f"q must be [N, {self.n_dofs}], got {tuple(q.shape)}")
        if q.device != self.device:
            raise RuntimeError(f"q is on {q.device} but the model lives on {self.device}")
        n = q.shape[0]
        pose = torch.empty((n, 12), dtype=torch.float32, device=q.device)
        lin = torch.empty((n, 3, self.n_dofs), dtype=torch.float32, device=q.device)
        ang = torch.empty_like(lin)
        check(self._L.trk_geometric_jacobian(self.handle, q.data_ptr(), n, int(link), int(col_mask) & 0xFFFFFFFF, pose.data_ptr(), lin.data_ptr(),
                                           ang.data_ptr(), stream_ptr(q.device)), "trk_geometric_jacobian", self._L)
        return pose, lin, ang

    def fk_backward(self, q: torch.Tensor, G: torch.Tensor, link_sel=None) -> torch.Tensor:
        q = require_cuda_f32(q, "q")
        G = require_cuda_f32(G, "grad_H")
        p, n_sel, keep = self._sel(link_sel)
        if tuple(G.shape) != (q.shape[0], n_sel, 4, 4):
            raise RuntimeError(f"grad_H must be [{q.shape[0]}, {n_sel}, 4, 4], got {tuple(G.shape)}")
        grad = torch.empty_like(q)
        check(self._L.trk_fk_backward(self.handle, q.data_ptr(), q.shape[0], p, n_sel, G.data_ptr(), grad.data_ptr(),
                                    stream_ptr(q.device)), "trk_fk_backward", self._L)
        return grad


    def ik_adam(self, q0: torch.Tensor, H_target, ee_link: int, lower, upper, n_iters: int, n_updates: Optional[int] = None,
                lr=1e-2, betas=(0.9, 0.999), eps=1e-8, w_se3=1.0, w_joint_limits=300.0, se3_eps=1e-1, want_loss_stats=False,
                best_k: int = 0):
        """Fused Adam IK (trk_ik_adam).  q0 [S, D] is not modified.  Returns dict(q, err, valid (bool), valid_count
        [n_iters] int32, loss_sum / loss_sqsum [n_iters] or None).  Compiled-in robots run the register-resident kernel,
        any other tree the runtime-tree one (state in shared-memory columns)."""
        q = require_cuda_f32(q0, "q0").clone()
        S, D = q.shape
        dev = q.device
        n_updates = n_iters if n_updates is None else n_updates
        # Adam's per-iteration bias corrections: device-resident, cached per schedule (two small uploads otherwise
        # cost more than the kernel for a few thousand seeds)
        key = (int(n_iters), float(lr), float(betas[0]), float(betas[1]), str(dev))
        sched = self._ik_schedules.get(key)
        if sched is None:
            t = np.arange(1, n_iters + 1, dtype=np.float64)
            sched = (torch.as_tensor((lr / (1.0 - betas[0] ** t)).astype(np.float32), device=dev),
                     torch.as_tensor(np.sqrt(1.0 - betas[1] ** t).astype(np.float32), device=dev))
            if len(self._ik_schedules) > 16:
                self._ik_schedules.clear()
            self._ik_schedules[key] = sched
        step_size, bc2 = sched
        lo = _c32(lower.detach().cpu().numpy() if torch.is_tensor(lower) else lower, np.float32).reshape(-1)
        hi = _c32(upper.detach().cpu().numpy() if torch.is_tensor(upper) else upper, np.float32).reshape(-1)
        if lo.shape[0] != D or hi.shape[0] != D:
            raise RuntimeError(f"lower/upper must have {D} entries")
        tgt_seed, tgt_host = None, None
        Ht = H_target.detach() if torch.is_tensor(H_target) else torch.as_tensor(np.asarray(H_target, np.float32))
        Ht = Ht.reshape(-1, 4, 4)
        if Ht.shape[0] == 1:
            tgt_host = _c32(Ht[0].cpu().numpy(), np.float32).reshape(16)
        elif Ht.shape[0] == S:
            tgt_seed = require_cuda_f32(Ht.to(dev), "H_target")
        else:
            raise RuntimeError("H_target must be [4,4], [1,4,4] or [S,4,4]")
        p = IKParamsC(int(ee_link), int(n_iters), int(n_updates), float(lr), float(betas[0]), float(betas[1]), float(eps),
                      float(w_se3), float(w_joint_limits), float(se3_eps), _fp(lo), _fp(hi), None if tgt_host is None else _fp(tgt_host))
        err = torch.empty(S, dtype=torch.float32, device=dev)
        valid = torch.empty(S, dtype=torch.uint8, device=dev)
        count = torch.empty(n_iters, dtype=torch.int32, device=dev)
        ls = torch.empty(n_iters, dtype=torch.float32, device=dev) if want_loss_stats else None
        lq = torch.empty(n_iters, dtype=torch.float32, device=dev) if want_loss_stats else None
        ptr = lambda x: None if x is None else x.data_ptr()
        best = None
        if best_k > 0 and not want_loss_stats:
            # per-block candidates for distributed.gather_best_solutions (trk_ik_adam_best); robots the fused selection does
            # not serve fall through to the plain call and the caller selects over all seeds
            best = torch.empty(((S + 127) // 128) * int(best_k), dtype=torch.int64, device=dev)
            rc = self._L.trk_ik_adam_best(self.handle, C.byref(p), step_size.data_ptr(), bc2.data_ptr(), ptr(tgt_seed), S, q.data_ptr(),
                                        err.data_ptr(), valid.data_ptr(), count.data_ptr(), best.data_ptr(), int(best_k), stream_ptr(dev))
            if rc == -4:      # TRK_E_UNSUPPORTED
                best = None
            else:
                check(rc, "trk_ik_adam_best", self._L)
        if best is None:
            check(self._L.trk_ik_adam(self.handle, C.byref(p), step_size.data_ptr(), bc2.data_ptr(), ptr(tgt_seed), S, q.data_ptr(),
                                    err.data_ptr(), valid.data_ptr(), count.data_ptr(), ptr(ls), ptr(lq), stream_ptr(dev)), "trk_ik_adam", self._L)
        return dict(q=q, err=err, valid=valid.bool(), valid_count=count, loss_sum=ls, loss_sqsum=lq, best_keys=best)


class DeviceCost:
    """Device-resident planning-objective tables (trk_cost) bound to a DeviceModel."""

    def __init__(self, model: DeviceModel, cost):
        cost.validate(model.n_links)
        self.model, self.cost = model, cost
        k = self._keep = dict(
            pt_link=_c32(cost.pt_link, np.int32), mobj=_c32(cost.pt_margin_obj, np.float32), mws=_c32(cost.pt_margin_ws, np.float32),
            obj_start=_c32(cost.obj_start, np.int32), obj_ident=_c32(cost.obj_identity, np.int32),
            obj_R=_c32(cost.obj_R, np.float32), obj_pos=_c32(cost.obj_pos, np.float32),
            sph_c=_c32(cost.sph_center, np.float32), sph_r=_c32(cost.sph_radius, np.float32),
            pi=_c32(cost.pair_i, np.int32), pj=_c32(cost.pair_j, np.int32), pm=_c32(cost.pair_margin, np.float32),
            obs=_c32(cost.obj_box_start, np.int32), bc=_c32(cost.box_center, np.float32), bh=_c32(cost.box_half, np.float32),
            br=_c32(cost.box_radius, np.float32), bf=_c32(cost.box_rounded, np.int32))
        cd = CostDesc()
        cd.n_points, cd.pt_link, cd.pt_margin_obj, cd.pt_margin_ws = cost.n_points, _ip(k["pt_link"]), _fp(k["mobj"]), _fp(k["mws"])
        cd.n_objects, cd.obj_start, cd.obj_identity = cost.n_objects, _ip(k["obj_start"]), _ip(k["obj_ident"])
        cd.obj_R, cd.obj_pos = _fp(k["obj_R"]), _fp(k["obj_pos"])
        cd.n_spheres, cd.sph_center, cd.sph_radius = cost.n_spheres, _fp(k["sph_c"]), _fp(k["sph_r"])
        cd.n_boxes, cd.obj_box_start, cd.box_center, cd.box_half = cost.n_boxes, _ip(k["obs"]), _fp(k["bc"]), _fp(k["bh"])
        cd.box_radius, cd.box_rounded = _fp(k["br"]), _ip(k["bf"])
        cd.grid_obj = int(getattr(cost, "grid_obj", -1))
        if cd.grid_obj >= 0:     # the voxel grid stays in HBM; the handle only records its address
            g = cost.grid_data
            g = g if torch.is_tensor(g) else torch.as_tensor(np.ascontiguousarray(g, np.float32))
            g = g.detach().to(device=model.device, dtype=torch.float32).contiguous()
            if tuple(g.shape) != (*[int(v) for v in cost.grid_n], 4):
                raise RuntimeError(f"voxel SDF data must be [n0, n1, n2, 4] = {tuple(cost.grid_n)} + (4,), got {tuple(g.shape)}")
            k["grid"] = g
            cd.grid_n = (C.c_int32 * 3)(*[int(v) for v in cost.grid_n])
            cd.grid_lo = (C.c_float * 3)(*[float(v) for v in cost.grid_lo])
            cd.grid_extent = (C.c_float * 3)(*[float(v) for v in cost.grid_extent])
            cd.grid_data = g.data_ptr()
        cd.ws_enabled = int(cost.ws_enabled)
        cd.ws_min = (C.c_float * 3)(*[float(v) for v in cost.ws_min])
        cd.ws_max = (C.c_float * 3)(*[float(v) for v in cost.ws_max])
        cd.n_pairs, cd.pair_i, cd.pair_j, cd.pair_margin = cost.n_pairs, _ip(k["pi"]), _ip(k["pj"]), _fp(k["pm"])
        cd.ee_enabled, cd.ee_link = int(cost.ee_enabled), int(cost.ee_link)
        cd.ee_w_pos, cd.ee_w_rot, cd.ee_square = float(cost.ee_w_pos), float(cost.ee_w_rot), int(cost.ee_square)
        cd.ee_target = (C.c_float * 16)(*[float(v) for v in np.asarray(cost.ee_target, np.float32).reshape(16)])
        cd.clamp_self, cd.clamp_obj, cd.clamp_ws = int(cost.clamp_self), int(cost.clamp_obj), int(cost.clamp_ws)
        cd.weights = (C.c_float * 4)(*[float(v) for v in cost.weights])
        h = C.c_void_p()
        check(model._L.trk_cost_create(model.handle, C.byref(cd), C.byref(h)), "trk_cost_create", self.model._L)
        self.handle = h

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and getattr(getattr(self, "model", None), "_L", None) is not None:
            try:
                self.model._L.trk_cost_destroy(h)
            except Exception:
                pass
            self.handle = None

    def set_ee_target(self, H) -> None:
        H = np.ascontiguousarray(np.asarray(H.detach().cpu().numpy() if torch.is_tensor(H) else H, np.float32).reshape(16))
        check(self.model._L.trk_cost_set_ee_target(self.handle, _fp(H), stream_ptr(self.model.device)), "trk_cost_set_ee_target", self.model._L)
        self.cost.ee_target = H.reshape(4, 4).copy()

    def set_weights(self, w) -> None:
        w = np.ascontiguousarray(np.asarray(w, np.float32).reshape(4))
        check(self.model._L.trk_cost_set_weights(self.handle, _fp(w), stream_ptr(self.model.device)), "trk_cost_set_weights", self.model._L)
        self.cost.weights = w.copy()

    def objective(self, q: torch.Tensor, want_grad=True, want_terms=True, want_occupancy=False, want_argmin=False,
                  ee_target_cfg: Optional[torch.Tensor] = None, margin_override: Optional[float] = None,
                  out_total: Optional[torch.Tensor] = None, out_grad: Optional[torch.Tensor] = None):
        """Raw fused call.  Returns dict(total, terms, grad_q, occupancy, argmin) (None where not requested).
        out_total [n] / out_grad [n, D]: optional preallocated contiguous fp32 CUDA outputs."""
        m = self.model
        q = require_cuda_f32(q, "q")
        if q.ndim != 2 or q.shape[1] != m.n_dofs:
            raise RuntimeError(f"q must be [N, {m.n_dofs}], got {tuple(q.shape)}")
        if q.device != m.device:
            raise RuntimeError(f"q is on {q.device} but the model lives on {m.device}")
        n = q.shape[0]
        dev = q.device
        total = torch.empty(n, dtype=torch.float32, device=dev) if out_total is None else out_total
        terms = torch.empty((n, 4), dtype=torch.float32, device=dev) if want_terms else None
        grad = (torch.empty_like(q) if out_grad is None else out_grad) if want_grad else None
        for name, buf, shape in (("out_total", out_total, (n,)), ("out_grad", out_grad, (n, m.n_dofs))):
            if buf is not None and (tuple(buf.shape) != shape or buf.dtype != torch.float32 or buf.device != dev
                                    or not buf.is_contiguous()):
                raise RuntimeError(f"{name} must be a contiguous float32 tensor of shape {shape} on {dev}")
        occ = torch.empty(n, dtype=torch.uint8, device=dev) if want_occupancy else None
        amin = torch.empty((n, self.cost.n_objects, self.cost.n_points), dtype=torch.int32, device=dev) if want_argmin else None
        tgt = None
        if ee_target_cfg is not None:
            tgt = require_cuda_f32(ee_target_cfg, "ee_target_cfg")
            if tuple(tgt.shape) != (n, 4, 4):
                raise RuntimeError(f"ee_target_cfg must be [{n}, 4, 4]")
        mo = float("nan") if margin_override is None else float(margin_override)
        ptr = lambda t: None if t is None else t.data_ptr()
        check(self.model._L.trk_objective(m.handle, self.handle, q.data_ptr(), n, ptr(tgt), mo, ptr(total), ptr(terms), ptr(grad),
                                  ptr(occ), ptr(amin), stream_ptr(dev)), "trk_objective", self.model._L)
        return dict(total=total, terms=terms, grad_q=grad, occupancy=occ, argmin=amin)

    MAX_BLOCK_SUMS = 8192

    def objective_sum(self, q: torch.Tensor, out_grad: torch.Tensor, out_total: Optional[torch.Tensor] = None,
                      out_block_sums: Optional[torch.Tensor] = None, margin_override: Optional[float] = None) -> torch.Tensor:
        """Cost + gradient with the first stage of the batch reduction fused into the kernel (trk_objective_sums).  Returns the
        per-block partial sums [MAX_BLOCK_SUMS] fp32 (zero-padded): `.sum(dtype=torch.float64)` is sum(cost), ready for
        distributed.total_cost_from_partials.  out_total is optional: planners that only need the scalar skip 4 B / configuration."""
        m = self.model
        q = require_cuda_f32(q, "q")
        n, dev = q.shape[0], q.device
        if q.ndim != 2 or q.shape[1] != m.n_dofs or dev != m.device:
            raise RuntimeError(f"q must be [N, {m.n_dofs}] on {m.device}")
        bs = torch.empty(self.MAX_BLOCK_SUMS, dtype=torch.float32, device=dev) if out_block_sums is None else out_block_sums
        for name, buf, shape in (("out_total", out_total, (n,)), ("out_grad", out_grad, (n, m.n_dofs)), ("out_block_sums", bs, (self.MAX_BLOCK_SUMS,))):
            if buf is not None and (tuple(buf.shape) != shape or buf.dtype != torch.float32 or buf.device != dev or not buf.is_contiguous()):
                raise RuntimeError(f"{name} must be a contiguous float32 tensor of shape {shape} on {dev}")
        mo = float("nan") if margin_override is None else float(margin_override)
        check(self.model._L.trk_objective_sums(m.handle, self.handle, q.data_ptr(), n, None, mo, None if out_total is None else out_total.data_ptr(),
                                       out_grad.data_ptr(), bs.data_ptr(), stream_ptr(dev)), "trk_objective_sums", self.model._L)
        return bs

    def objective_host(self, q_host: torch.Tensor, cost_host: torch.Tensor, grad_host: Optional[torch.Tensor] = None,
                       n_chunks: int = 8) -> None:
        """trk_objective_host: contiguous fp32 HOST tensors in and out (pin them for overlapped copies); returns
        when cost_host / grad_host are filled."""
        for t, name in ((q_host, "q_host"), (cost_host, "cost_host"), (grad_host, "grad_host")):
            if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError(f"{name} must be a contiguous float32 host tensor")
        n = q_host.shape[0]
        if q_host.ndim != 2 or q_host.shape[1] != self.model.n_dofs or cost_host.shape != (n,) or \
                (grad_host is not None and tuple(grad_host.shape) != (n, self.model.n_dofs)):
            raise RuntimeError("objective_host: shape mismatch")
        check(self.model._L.trk_objective_host(self.model.handle, self.handle, q_host.data_ptr(), n, int(n_chunks), cost_host.data_ptr(),
                                       None if grad_host is None else grad_host.data_ptr()), "trk_objective_host", self.model._L)

    def cost_stage(self, link_pos: torch.Tensor, ee_pose: Optional[torch.Tensor] = None, want_grad=True,
                   want_occupancy=False, want_argmin=False, ee_target_cfg=None, margin_override: Optional[float] = None):
        m = self.model
        link_pos = require_cuda_f32(link_pos, "link_pos")
        if link_pos.ndim != 3 or link_pos.shape[1:] != (m.n_links, 3):
            raise RuntimeError(f"link_pos must be [N, {m.n_links}, 3], got {tuple(link_pos.shape)}")
        n, dev = link_pos.shape[0], link_pos.device
        if dev != m.device:
            raise RuntimeError(f"link_pos is on {dev}, the model on {m.device}")
        if ee_pose is not None:
            ee_pose = require_cuda_f32(ee_pose, "ee_pose")
            if tuple(ee_pose.shape) != (n, 4, 4) or ee_pose.device != dev:
                raise RuntimeError(f"ee_pose must be [{n}, 4, 4] on {dev}, got {tuple(ee_pose.shape)} on {ee_pose.device}")
        total = torch.empty(n, dtype=torch.float32, device=dev)
        terms = torch.empty((n, 4), dtype=torch.float32, device=dev)
        gpos = torch.empty_like(link_pos) if want_grad else None
        gee = torch.zeros((n, 4, 4), dtype=torch.float32, device=dev) if (want_grad and ee_pose is not None) else None
        occ = torch.empty(n, dtype=torch.uint8, device=dev) if want_occupancy else None
        amin = torch.empty((n, self.cost.n_objects, self.cost.n_points), dtype=torch.int32, device=dev) if want_argmin else None
        tgt = None if ee_target_cfg is None else require_cuda_f32(ee_target_cfg, "ee_target_cfg")
        if tgt is not None and (tuple(tgt.shape) != (n, 4, 4) or tgt.device != dev):
            raise RuntimeError(f"ee_target_cfg must be [{n}, 4, 4] on {dev}, got {tuple(tgt.shape)} on {tgt.device}")
        mo = float("nan") if margin_override is None else float(margin_override)
        ptr = lambda t: None if t is None else t.data_ptr()
        check(self.model._L.trk_cost_stage(m.handle, self.handle, link_pos.data_ptr(), ptr(ee_pose), n, ptr(tgt), mo, ptr(total),
                                   ptr(terms), ptr(gpos), ptr(gee), ptr(occ), ptr(amin), stream_ptr(dev)), "trk_cost_stage", self.model._L)
        return dict(total=total, terms=terms, grad_pos=gpos, grad_ee=gee, occupancy=occ, argmin=amin)
