"""ctypes loader for oracle/trk_oracle.c (TEST INFRASTRUCTURE ONLY - see the header of that file).

Used by tests/ as a fast bit-exact checker and by bench.py as the CPU port that is timed on the
host cores (`cpu_baseline`, `--impl reference`).  Never imported by torch_robotics_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libtrk_oracle.so")
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)


class OrcTree(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("n_dofs", C.c_int32), ("parent", ip), ("kind", ip), ("axis_k", ip),
                ("axis_s", fp), ("axis", fp), ("q_index", ip), ("R_fixed", fp), ("trans", fp), ("lo", fp), ("hi", fp),
                ("has_lim", ip), ("topo", ip), ("base_pose", fp)]


class OrcCost(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("pt_link", ip), ("pt_margin_obj", fp), ("pt_margin_ws", fp),
                ("n_objects", C.c_int32), ("obj_start", ip), ("obj_identity", ip), ("obj_R", fp), ("obj_pos", fp),
                ("n_spheres", C.c_int32), ("sph_center", fp), ("sph_radius", fp),
                ("n_boxes", C.c_int32), ("obj_box_start", ip), ("box_center", fp), ("box_half", fp), ("box_radius", fp), ("box_rounded", ip),
                ("grid_obj", C.c_int32), ("grid_n", C.c_int32 * 3), ("grid_lo", C.c_float * 3), ("grid_extent", C.c_float * 3), ("grid_data", fp),
                ("ws_enabled", C.c_int32), ("ws_min", C.c_float * 3), ("ws_max", C.c_float * 3),
                ("n_pairs", C.c_int32), ("pair_i", ip), ("pair_j", ip), ("pair_margin", fp),
                ("ee_enabled", C.c_int32), ("ee_link", C.c_int32), ("ee_w_pos", C.c_float), ("ee_w_rot", C.c_float),
                ("ee_square", C.c_int32), ("ee_target", C.c_float * 16),
                ("clamp_self", C.c_int32), ("clamp_obj", C.c_int32), ("clamp_ws", C.c_int32), ("weights", C.c_float * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = C.CDLL(LIB)
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        vp = C.c_void_p
        L.orc_fk_forward.argtypes = [C.POINTER(OrcTree), vp, C.c_int64, vp, C.c_int32, vp]
        L.orc_fk_backward.argtypes = [C.POINTER(OrcTree), vp, C.c_int64, vp, C.c_int32, vp, vp]
        L.orc_objective.argtypes = [C.POINTER(OrcTree), C.POINTER(OrcCost), vp, C.c_int64, vp, C.c_float, vp, vp, vp, vp, vp]
        for f in (L.orc_fk_forward, L.orc_fk_backward, L.orc_objective, L.orc_set_threads):
            f.restype = None
        _lib = L
    return _lib


def _a(x, dt):
    return np.ascontiguousarray(np.asarray(x, dt))


class Tree:
    def __init__(self, d):
        k = self.keep = dict(parent=_a(d.parent, np.int32), kind=_a(d.kind, np.int32), axis_k=_a(d.axis_k, np.int32),
                             axis_s=_a(d.axis_s, np.float32), axis=_a(d.axis, np.float32), q_index=_a(d.q_index, np.int32),
                             R_fixed=_a(d.R_fixed, np.float32), trans=_a(d.trans, np.float32), lo=_a(d.lo, np.float32),
                             hi=_a(d.hi, np.float32), has_lim=_a(d.has_lim, np.int32), topo=_a(d.topo, np.int32))
        P = lambda n: k[n].ctypes.data_as(ip if k[n].dtype == np.int32 else fp)
        base = None
        bR, bt = np.asarray(getattr(d, "base_R", np.eye(3)), np.float32), np.asarray(getattr(d, "base_t", np.zeros(3)), np.float32)
        if not np.array_equal(bR, np.eye(3, dtype=np.float32)) or np.any(bt != 0):      # root pose = base pose (robot_tree.py:133)
            k["base"] = _a(np.concatenate([bR.reshape(9), bt.reshape(3)]), np.float32)
            base = k["base"].ctypes.data_as(fp)
        self.s = OrcTree(d.n_links, d.n_dofs, P("parent"), P("kind"), P("axis_k"), P("axis_s"), P("axis"), P("q_index"),
                         P("R_fixed"), P("trans"), P("lo"), P("hi"), P("has_lim"), P("topo"), base)
        self.L, self.D = d.n_links, d.n_dofs


class Cost:
    def __init__(self, c):
        k = self.keep = dict(pt_link=_a(c.pt_link, np.int32), mobj=_a(c.pt_margin_obj, np.float32), mws=_a(c.pt_margin_ws, np.float32),
                             obj_start=_a(c.obj_start, np.int32), obj_ident=_a(c.obj_identity, np.int32), obj_R=_a(c.obj_R, np.float32),
                             obj_pos=_a(c.obj_pos, np.float32), sc=_a(c.sph_center, np.float32), sr=_a(c.sph_radius, np.float32),
                             pi=_a(c.pair_i, np.int32), pj=_a(c.pair_j, np.int32), pm=_a(c.pair_margin, np.float32),
                             obs=_a(c.obj_box_start, np.int32), bc=_a(c.box_center, np.float32), bh=_a(c.box_half, np.float32),
                             br=_a(c.box_radius, np.float32), bf=_a(c.box_rounded, np.int32))
        P = lambda n: k[n].ctypes.data_as(ip if k[n].dtype == np.int32 else fp)
        s = self.s = OrcCost()
        s.n_points, s.pt_link, s.pt_margin_obj, s.pt_margin_ws = c.n_points, P("pt_link"), P("mobj"), P("mws")
        s.n_objects, s.obj_start, s.obj_identity, s.obj_R, s.obj_pos = c.n_objects, P("obj_start"), P("obj_ident"), P("obj_R"), P("obj_pos")
        s.n_spheres, s.sph_center, s.sph_radius = c.n_spheres, P("sc"), P("sr")
        s.n_boxes, s.obj_box_start, s.box_center, s.box_half = c.n_boxes, P("obs"), P("bc"), P("bh")
        s.box_radius, s.box_rounded = P("br"), P("bf")
        s.grid_obj = int(getattr(c, "grid_obj", -1))
        if s.grid_obj >= 0:
            g = c.grid_data
            k["grid"] = _a(g if isinstance(g, np.ndarray) else g.detach().cpu().numpy(), np.float32)
            s.grid_n = (C.c_int32 * 3)(*[int(v) for v in c.grid_n])
            s.grid_lo = (C.c_float * 3)(*[float(v) for v in c.grid_lo])
            s.grid_extent = (C.c_float * 3)(*[float(v) for v in c.grid_extent])
            s.grid_data = k["grid"].ctypes.data_as(fp)
        s.ws_enabled = int(c.ws_enabled)
        s.ws_min = (C.c_float * 3)(*[float(v) for v in c.ws_min])
        s.ws_max = (C.c_float * 3)(*[float(v) for v in c.ws_max])
        s.n_pairs, s.pair_i, s.pair_j, s.pair_margin = c.n_pairs, P("pi"), P("pj"), P("pm")
        s.ee_enabled, s.ee_link, s.ee_w_pos, s.ee_w_rot, s.ee_square = int(c.ee_enabled), int(c.ee_link), float(c.ee_w_pos), float(c.ee_w_rot), int(c.ee_square)
        s.ee_target = (C.c_float * 16)(*[float(v) for v in np.asarray(c.ee_target, np.float32).reshape(16)])
        s.clamp_self, s.clamp_obj, s.clamp_ws = int(c.clamp_self), int(c.clamp_obj), int(c.clamp_ws)
        s.weights = (C.c_float * 4)(*[float(v) for v in c.weights])
        self.S, self.nObj = c.n_points, c.n_objects


def fk_forward(tree: Tree, q, sel=None):
    q = _a(q, np.float32)
    n = q.shape[0]
    sel_a = None if sel is None else _a(sel, np.int32)
    n_sel = tree.L if sel is None else len(sel_a)
    H = np.empty((n, n_sel, 4, 4), np.float32)
    lib().orc_fk_forward(C.byref(tree.s), q.ctypes.data, n, None if sel_a is None else sel_a.ctypes.data, n_sel, H.ctypes.data)
    return H


def fk_backward(tree: Tree, q, G, sel=None):
    q, G = _a(q, np.float32), _a(G, np.float32)
    n = q.shape[0]
    sel_a = None if sel is None else _a(sel, np.int32)
    n_sel = tree.L if sel is None else len(sel_a)
    grad = np.empty_like(q)
    lib().orc_fk_backward(C.byref(tree.s), q.ctypes.data, n, None if sel_a is None else sel_a.ctypes.data, n_sel, G.ctypes.data, grad.ctypes.data)
    return grad


def objective(tree: Tree, cost: Cost, q, ee_target_cfg=None, margin_override=None, want_grad=True, out=None):
    q = _a(q, np.float32)
    n = q.shape[0]
    if out is None:
        out = dict(total=np.empty(n, np.float32), terms=np.empty((n, 4), np.float32),
                   grad_q=np.empty_like(q) if want_grad else None, occ=np.empty(n, np.uint8),
                   argmin=np.empty((n, cost.nObj, cost.S), np.int32))
    tgt = None if ee_target_cfg is None else _a(ee_target_cfg, np.float32)
    mo = float("nan") if margin_override is None else float(margin_override)
    P = lambda a: None if a is None else a.ctypes.data
    lib().orc_objective(C.byref(tree.s), C.byref(cost.s), q.ctypes.data, n, P(tgt), mo, P(out["total"]), P(out["terms"]),
                        P(out.get("grad_q")), P(out.get("occ")), P(out.get("argmin")))
    return out


def max_threads():
    return lib().orc_max_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))
