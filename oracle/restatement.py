"""CPU restatement (numpy) of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline/reference leg may import
this module; the product path (torch_robotics_b200/) never does and fails loudly without its
CUDA library.

What is restated (all citations relative to /root/reference/torch_robotics):
  * per-link joint transform + chain composition   torch_kinematics_tree/models/rigid_body.py:146-211,
                                                   geometrics/spatial_vector.py:8-47, geometrics/utils.py:11-17
  * homogeneous output                             models/robot_tree.py:267-301, geometrics/frame.py:81-85
  * self-collision / object / workspace costs      torch_planning_objectives/fields/distance_fields.py:107-130,
                                                   :194-215, :283-291, :307-332; environments/primitives.py:108-112,:387-405
  * EE SE(3) distance                              geometrics/utils.py:130-154,:209-224; distance_fields.py:347-356
  * IK loss / termination / Adam loop              models/robot_tree.py:303-442
  * boxes / rounded boxes, posed objects           environments/primitives.py:198-334, :387-405
  * voxel SDF lookup with surrogate gradient       environments/grid_map_sdf.py:96-112
  * link Jacobians (position + quaternion)         models/robot_tree.py:250-265, geometrics/quaternion.py:135-166
  * reverse mode: the reference relies on autograd; the analytic form restated here is the one
    SURVEY.md section 8(a) derives (per-link wrench accumulated leaf->root).

Pinned against: tests/golden/*.npz, produced by running the unmodified reference in the build
container (tests/golden/make_golden.py).  The reference ships no tests/golden vectors of its own.

dtype: np.float32 reproduces the reference's fp32 arithmetic; decision values (distances that feed
`<` comparisons and arg-min/arg-max) follow ONE canonical operation order, shared with
oracle/trk_oracle.c and the CUDA kernels so that masks and indices can be compared bit-exactly:
    |v| = sqrt(fma(vz, vz, fma(vy, vy, vx*vx)))        (what torch.linalg.norm does on CPU; verified)
    sdf = |x - c| - r ;  first index wins ties in min/max.
np.float64 is available for finite-difference style checks of the analytic backward.
"""
from __future__ import annotations

import numpy as np

KIND_FIXED, KIND_REVOLUTE, KIND_PRISMATIC = 0, 1, 2


# ----------------------------------------------------------------------------------------------
# exact fp32 fused multiply-add in numpy
# ----------------------------------------------------------------------------------------------
def fma32(a, b, c):
    """round_to_nearest_even(a*b + c) for float32 arrays, bit-exact with C fmaf().

    a*b is exact in float64 (24+24 bits).  p + c is formed in float64 with its rounding error
    recovered by TwoSum; the only way the subsequent float64->float32 rounding can go wrong
    (double rounding) is when the float64 sum sits exactly on a float32 tie, which is repaired
    using the sign of the recovered error."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    c = np.asarray(c, np.float32)
    p = a.astype(np.float64) * b.astype(np.float64)
    c64 = c.astype(np.float64)
    s = p + c64
    bb = s - p
    e = (p - (s - bb)) + (c64 - bb)          # exact: s + e == p + c
    r = s.astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        up = np.nextafter(r, np.float32(np.inf)).astype(np.float64)
        dn = np.nextafter(r, np.float32(-np.inf)).astype(np.float64)
        r64 = r.astype(np.float64)
        tie_up = (s == (r64 + up) * 0.5) & (e > 0)     # s rounded down to even r, true value above the tie
        tie_dn = (s == (r64 + dn) * 0.5) & (e < 0)
    r = np.where(tie_up, up.astype(np.float32), r)
    r = np.where(tie_dn, dn.astype(np.float32), r)
    return r.astype(np.float32)


def norm3(vx, vy, vz, dtype):
    """Canonical |v| (see module docstring)."""
    if dtype == np.float32:
        s = (vx * vx).astype(np.float32)
        s = fma32(vy, vy, s)
        s = fma32(vz, vz, s)
        return np.sqrt(s).astype(np.float32)
    return np.sqrt(vx * vx + vy * vy + vz * vz)


# ----------------------------------------------------------------------------------------------
# forward kinematics
# ----------------------------------------------------------------------------------------------
def _axis_rot(k, ang):
    """spatial_vector.py:8-47 - [N,3,3] rotation about axis k."""
    c, s = np.cos(ang), np.sin(ang)
    R = np.zeros(ang.shape + (3, 3), ang.dtype)
    i, j = [(1, 2), (2, 0), (0, 1)][k]
    R[..., k, k] = 1
    R[..., i, i] = c
    R[..., j, j] = c
    R[..., i, j] = -s
    R[..., j, i] = s
    return R


def joint_angles(desc, q, dtype):
    """Clamped joint value per body and the pass-through mask of torch.clamp's backward
    (rigid_body.py:157-160; clamp passes gradient where lo <= q <= hi)."""
    q = np.asarray(q, dtype)
    N = q.shape[0]
    L = desc.n_links
    theta = np.zeros((N, L), dtype)
    mask = np.zeros((N, L), dtype)
    for l in range(L):
        qi = int(desc.q_index[l])
        if qi < 0:
            continue
        v = q[:, qi]
        if desc.has_lim[l]:
            lo, hi = dtype(np.float32(desc.lo[l])), dtype(np.float32(desc.hi[l]))
            theta[:, l] = np.minimum(np.maximum(v, lo), hi)
            mask[:, l] = ((v >= lo) & (v <= hi)).astype(dtype)
        else:
            theta[:, l] = v
            mask[:, l] = 1
    return theta, mask


def fk_forward(desc, q, dtype=np.float32, base_R=None, base_t=None):
    """World rotation [N,L,3,3] and position [N,L,3] of every body, body (= <link>) order.

    rigid_body.py:146-211 composes leaf->root; the product is the same up to fp32 re-association:
    R_w[l] = R_w[parent] @ (R_fixed @ Rot_k(s*theta)),  p_w[l] = R_w[parent] @ t_l + p_w[parent]."""
    q = np.asarray(q, dtype)
    if q.ndim == 1:
        q = q[None]
    N, L = q.shape[0], desc.n_links
    theta, _ = joint_angles(desc, q, dtype)
    Rw = np.zeros((N, L, 3, 3), dtype)
    pw = np.zeros((N, L, 3), dtype)
    bR = np.eye(3, dtype=dtype) if base_R is None else np.asarray(base_R, dtype)
    bt = np.zeros(3, dtype) if base_t is None else np.asarray(base_t, dtype)
    for l in [int(x) for x in desc.topo]:
        p = int(desc.parent[l])
        if p < 0:
            Rw[:, l] = bR            # rigid_body.py:201-208: the root reports the base pose
            pw[:, l] = bt
            continue
        Rf = desc.R_fixed[l].astype(dtype)
        t = np.broadcast_to(desc.trans[l].astype(dtype), (N, 3))
        kind = int(desc.kind[l])
        if kind == KIND_REVOLUTE:
            ang = (dtype(desc.axis_s[l]) * theta[:, l]).astype(dtype)
            Rl = (Rf[None] @ _axis_rot(int(desc.axis_k[l]), ang)).astype(dtype)    # :170-171
        else:
            Rl = np.broadcast_to(Rf, (N, 3, 3))
            if kind == KIND_PRISMATIC:
                t = (t + desc.axis[l].astype(dtype)[None] * theta[:, l:l + 1]).astype(dtype)   # :176-181
        Rw[:, l] = (Rw[:, p] @ Rl).astype(dtype)                                   # geometrics/utils.py:14
        pw[:, l] = ((Rw[:, p] @ t[..., None])[..., 0] + pw[:, p]).astype(dtype)    # :15
    return Rw, pw


def homogeneous(Rw, pw, sel=None):
    """frame.py:81-85 + robot_tree.py:71-72 - [N, n_sel, 4, 4]."""
    if sel is not None:
        Rw, pw = Rw[:, sel], pw[:, sel]
    H = np.zeros(Rw.shape[:2] + (4, 4), Rw.dtype)
    H[..., :3, :3] = Rw
    H[..., :3, 3] = pw
    H[..., 3, 3] = 1
    return H


# ----------------------------------------------------------------------------------------------
# reverse mode through the tree
# ----------------------------------------------------------------------------------------------
def _axial(M):
    return np.stack([M[..., 2, 1] - M[..., 1, 2], M[..., 0, 2] - M[..., 2, 0], M[..., 1, 0] - M[..., 0, 1]], -1)


def tree_backward(desc, q, Rw, pw, gR, gp, dtype=np.float32):
    """grad_q [N,D] from upstream gradients on every body's world rotation gR [N,L,3,3] and
    position gp [N,L,3].

    Each body contributes a force f = gp and a torque about the world origin
    n = p_w x gp + axial(gR R_w^T); subtree sums (F, N) are accumulated leaf->root, then
      revolute : dq = mask * (s * R_w[:, :, k]) . (N - p_w x F)
      prismatic: dq = mask * (R_w[parent] a) . F
    (equivalent to autograd through rigid_body.py:146-211)."""
    N_, L = Rw.shape[0], desc.n_links
    _, mask = joint_angles(desc, q, dtype)
    F = gp.astype(dtype).copy()
    T = (np.cross(pw, gp) + _axial(gR @ np.swapaxes(Rw, -1, -2))).astype(dtype)
    grad = np.zeros((N_, desc.n_dofs), dtype)
    for l in [int(x) for x in desc.topo[::-1]]:
        p = int(desc.parent[l])
        if p < 0:
            continue
        qi, kind = int(desc.q_index[l]), int(desc.kind[l])
        if kind == KIND_REVOLUTE:
            ax = dtype(desc.axis_s[l]) * Rw[:, l, :, int(desc.axis_k[l])]
            grad[:, qi] = mask[:, l] * np.sum(ax * (T[:, l] - np.cross(pw[:, l], F[:, l])), -1)
        elif kind == KIND_PRISMATIC:
            ax = (Rw[:, p] @ desc.axis[l].astype(dtype))
            grad[:, qi] = mask[:, l] * np.sum(ax * F[:, l], -1)
        F[:, p] += F[:, l]
        T[:, p] += T[:, l]
    return grad


def fk_backward(desc, q, G, sel=None, dtype=np.float32):
    """grad_q for a dense upstream gradient G [N, n_sel, 4, 4] on the homogeneous output
    (bottom row carries no gradient: frame.py:82 starts from a constant eye(4))."""
    q = np.asarray(q, dtype)
    Rw, pw = fk_forward(desc, q, dtype)
    L = desc.n_links
    sel = list(range(L)) if sel is None else list(sel)
    gR = np.zeros_like(Rw)
    gp = np.zeros_like(pw)
    G = np.asarray(G, dtype)
    for s, l in enumerate(sel):
        gR[:, l] += G[:, s, :3, :3]
        gp[:, l] += G[:, s, :3, 3]
    return tree_backward(desc, q, Rw, pw, gR, gp, dtype)


# ----------------------------------------------------------------------------------------------
# planning objective on link positions
# ----------------------------------------------------------------------------------------------
def rotation_to_quaternion(R, dtype=np.float32):
    """(w, x, y, z) of rotation matrices R [..., 3, 3] as geometrics/quaternion.py:135-166 picks it: of the four
    candidates, the one whose own component sqrt(1 +- m00 +- m11 +- m22) is largest (first on ties), so that
    component is positive."""
    R = np.asarray(R, dtype)
    m = [[R[..., i, j] for j in range(3)] for i in range(3)]
    t = np.stack([1 + m[0][0] + m[1][1] + m[2][2], 1 + m[0][0] - m[1][1] - m[2][2],
                  1 - m[0][0] + m[1][1] - m[2][2], 1 - m[0][0] - m[1][1] + m[2][2]], -1).astype(dtype)
    qa = np.sqrt(np.maximum(t, 0)).astype(dtype)
    rows = np.stack([
        np.stack([qa[..., 0] ** 2, m[2][1] - m[1][2], m[0][2] - m[2][0], m[1][0] - m[0][1]], -1),
        np.stack([m[2][1] - m[1][2], qa[..., 1] ** 2, m[1][0] + m[0][1], m[0][2] + m[2][0]], -1),
        np.stack([m[0][2] - m[2][0], m[1][0] + m[0][1], qa[..., 2] ** 2, m[1][2] + m[2][1]], -1),
        np.stack([m[1][0] - m[0][1], m[2][0] + m[0][2], m[2][1] + m[1][2], qa[..., 3] ** 2], -1)], -2).astype(dtype)
    cand = rows / (2 * np.maximum(qa[..., None], dtype(0.1)))
    i = np.argmax(qa, -1)
    return np.take_along_axis(cand, i[..., None, None], -2)[..., 0, :].astype(dtype)


def link_jacobian(desc, q, dtype=np.float32):
    """J [N, L, 7, D]: d (position xyz, quaternion wxyz) of every body / d q - what
    DifferentiableTree.compute_analytical_jacobian_all_links (robot_tree.py:250-265) obtains by autograd through FK and
    rotation_matrix_to_q.  For joint j on the path root -> l with world axis a_j and origin p_j:
      revolute : d p_l = a_j x (p_l - p_j),  d quat_l = 1/2 (0, a_j) (x) quat_l     (world-frame angular velocity)
      prismatic: d p_l = R_w[parent(j)] a_raw, d quat_l = 0
    times the clamp mask of joint j; all other columns are zero."""
    q = np.asarray(q, dtype)
    if q.ndim == 1:
        q = q[None]
    N, L, D = q.shape[0], desc.n_links, desc.n_dofs
    Rw, pw = fk_forward(desc, q, dtype)
    _, mask = joint_angles(desc, q, dtype)
    quat = rotation_to_quaternion(Rw, dtype)
    J = np.zeros((N, L, 7, D), dtype)
    for l in range(L):
        w, v = quat[:, l, 0], quat[:, l, 1:]
        b = l
        while b >= 0 and desc.parent[b] >= 0:
            qi, kind = int(desc.q_index[b]), int(desc.kind[b])
            if kind == KIND_REVOLUTE:
                a = dtype(desc.axis_s[b]) * Rw[:, b, :, int(desc.axis_k[b])]
                J[:, l, 0:3, qi] = mask[:, b, None] * np.cross(a, pw[:, l] - pw[:, b])
                J[:, l, 3, qi] = mask[:, b] * dtype(-0.5) * np.sum(a * v, -1)
                J[:, l, 4:7, qi] = mask[:, b, None] * dtype(0.5) * (w[:, None] * a + np.cross(a, v))
            elif kind == KIND_PRISMATIC:
                J[:, l, 0:3, qi] = mask[:, b, None] * (Rw[:, int(desc.parent[b])] @ desc.axis[b].astype(dtype))
            b = int(desc.parent[b])
    return J


def _box_sdf_and_grad(cost, kb, xo, dtype):
    """One box primitive at points xo [N,S,3] (object frame): signed distance and its gradient.
    primitives.py:219-223 (sharp: max_c(|x-c| - h)), :326-334 (rounded: min(max q, 0) + |relu(q)| - r,
    q = |x-c| - h + r).  Gradient = autograd of those expressions (first maximal component gets the max's share)."""
    c, hh, r = cost.box_center[kb].astype(dtype), cost.box_half[kb].astype(dtype), dtype(cost.box_radius[kb])
    v = (xo - c).astype(dtype)
    q = (np.abs(v) - hh).astype(dtype)
    sgn = np.sign(v).astype(dtype)
    w = np.zeros_like(q)
    if cost.box_rounded[kb]:
        q = (q + r).astype(dtype)
        mq = q.max(-1)
        rq = np.maximum(q, 0)
        nrm = norm3(rq[..., 0], rq[..., 1], rq[..., 2], dtype)
        sd = ((np.minimum(mq, 0) + nrm).astype(dtype) - r).astype(dtype)
        with np.errstate(invalid="ignore", divide="ignore"):
            w = np.where(nrm[..., None] > 0, rq / nrm[..., None], 0).astype(dtype)
        max_active = mq < 0
    else:
        mq = q.max(-1)
        sd = mq.astype(dtype)
        max_active = np.ones(mq.shape, bool)
    amax = np.argmax(q, -1)
    onehot = (np.arange(3)[None, None, :] == amax[..., None]) & max_active[..., None]
    return sd, (sgn * (w + onehot.astype(dtype))).astype(dtype)


def grid_cells(cost, x):
    """Cell index of points x [..., 3]: clamp(int(floor((x - lo) / extent * n)), 0, n - 1), every step rounded to fp32
    (grid_map_sdf.py:96-101: the grid shape enters as an int64 tensor that type promotion turns into fp32)."""
    x = np.asarray(x, np.float32)
    t = ((x - cost.grid_lo.astype(np.float32)) / cost.grid_extent.astype(np.float32)).astype(np.float32)
    t = np.floor((t * cost.grid_n.astype(np.float32)).astype(np.float32))
    with np.errstate(invalid="ignore"):
        i = np.where(np.isfinite(t), np.clip(t, -1.0, 2.0 ** 30), -1.0).astype(np.int64)
    return np.clip(i, 0, cost.grid_n.astype(np.int64) - 1)


def _grid_host(cost):
    g = cost.grid_data
    if not isinstance(g, np.ndarray):
        g = g.detach().cpu().numpy()
    return np.asarray(g, np.float32)


def object_term(cost, x, dtype=np.float32, margin_override=None):
    """Object-collision term on collision points x [N,S,3].

    primitives.py:387-405 (object frame), :108-112 (min_k |x-c_k| - r_k), distance_fields.py:316
    (stack objects), :113-124 (margin - sdf, max over objects, sum over points), :283-291 (occupancy).
    Returns dict(cost[N], gx[N,S,3], occ[N] bool, argmin[N,n_obj,S] int32, sdf[N,n_obj,S])."""
    N, S = x.shape[:2]
    nO = cost.n_objects
    sdf = np.zeros((N, nO, S), dtype)
    amin = np.zeros((N, nO, S), np.int32)
    gdir = np.zeros((N, nO, S, 3), dtype)      # d sdf / d x (world frame)
    for o in range(nO):
        k0, k1 = int(cost.obj_start[o]), int(cost.obj_start[o + 1])
        if cost.obj_identity[o]:
            xo = x
        else:
            Ro = cost.obj_R[o].astype(dtype)
            d = (x - cost.obj_pos[o].astype(dtype)).astype(dtype)
            if dtype == np.float32:
                xo = np.stack([fma32(d[..., 2], Ro[2, j], fma32(d[..., 1], Ro[1, j], d[..., 0] * Ro[0, j]))
                               for j in range(3)], -1)
            else:
                xo = d @ Ro
        best = np.full((N, S), np.inf, dtype)
        bidx = np.zeros((N, S), np.int32)
        bdir = np.zeros((N, S, 3), dtype)
        for k in range(k0, k1):
            v = (xo - cost.sph_center[k].astype(dtype)).astype(dtype)
            dist = norm3(v[..., 0], v[..., 1], v[..., 2], dtype)
            sd = (dist - dtype(cost.sph_radius[k])).astype(dtype)
            take = sd < best                      # strict: first index wins ties (torch.min)
            best = np.where(take, sd, best)
            bidx = np.where(take, k - k0, bidx)
            with np.errstate(invalid="ignore", divide="ignore"):
                dirk = np.where(dist[..., None] > 0, v / dist[..., None], 0).astype(dtype)
            bdir = np.where(take[..., None], dirk, bdir)
        kb0, kb1 = (int(cost.obj_box_start[o]), int(cost.obj_box_start[o + 1])) if getattr(cost, "n_boxes", 0) else (0, 0)
        for kb in range(kb0, kb1):   # boxes come after the object's spheres; strict < keeps the earlier primitive
            sd, gd = _box_sdf_and_grad(cost, kb, xo, dtype)
            take = sd < best
            best = np.where(take, sd, best)
            bidx = np.where(take, (k1 - k0) + (kb - kb0), bidx)
            bdir = np.where(take[..., None], gd, bdir)
        if o == getattr(cost, "grid_obj", -1):   # voxel SDF lookup, grid_map_sdf.py:96-112 (index always in fp32)
            ci = grid_cells(cost, xo)
            n1, n2 = int(cost.grid_n[1]), int(cost.grid_n[2])
            cell = _grid_host(cost)[ci[..., 0], ci[..., 1], ci[..., 2]]
            best, bdir = cell[..., 0].astype(dtype), cell[..., 1:4].astype(dtype)
            bidx = ((ci[..., 0] * n1 + ci[..., 1]) * n2 + ci[..., 2]).astype(np.int32)
        sdf[:, o], amin[:, o] = best, bidx
        gdir[:, o] = bdir if cost.obj_identity[o] else (bdir @ cost.obj_R[o].astype(dtype).T)
    margin = cost.pt_margin_obj.astype(dtype)[None, None, :]
    occ_margin = margin if margin_override is None else dtype(margin_override)
    occ = (sdf < occ_margin).any(-1).any(-1) if nO else np.zeros(N, bool)
    v = (-(sdf - margin)).astype(dtype)                                    # :116
    active = np.ones_like(v)
    if cost.clamp_obj:
        active = (v > 0).astype(dtype)
        v = np.maximum(v, 0)
    if nO == 0:
        return dict(cost=np.zeros(N, dtype), gx=np.zeros((N, S, 3), dtype), occ=occ, argmin=amin, sdf=sdf)
    omax = np.argmax(v, axis=1)                                            # first max wins (:122)
    vmax = np.take_along_axis(v, omax[:, None, :], 1)[:, 0]
    c = np.zeros(N, dtype)
    for s in range(S):
        c = (c + vmax[:, s]).astype(dtype)                                 # :124
    gsel = np.take_along_axis(gdir, omax[:, None, :, None], 1)[:, 0]
    asel = np.take_along_axis(active, omax[:, None, :], 1)[:, 0]
    gx = (-gsel * asel[..., None]).astype(dtype)
    return dict(cost=c, gx=gx, occ=occ, argmin=amin, sdf=sdf, omax=omax.astype(np.int32))


def ws_term(cost, x, dtype=np.float32, margin_override=None):
    """Workspace-boundary term (distance_fields.py:326-332, then :113-124 / :283-291)."""
    N, S = x.shape[:2]
    lo, hi = cost.ws_min.astype(dtype), cost.ws_max.astype(dtype)
    d = np.concatenate([(x - lo).astype(dtype), (hi - x).astype(dtype)], -1)      # [N,S,6]
    d = np.swapaxes(d, -1, -2)                                                    # [N,6,S]
    margin = cost.pt_margin_ws.astype(dtype)[None, None, :]
    occ_margin = margin if margin_override is None else dtype(margin_override)
    occ = (d < occ_margin).any(-1).any(-1)
    v = (-(d - margin)).astype(dtype)
    active = np.ones_like(v)
    if cost.clamp_ws:
        active = (v > 0).astype(dtype)
        v = np.maximum(v, 0)
    imax = np.argmax(v, axis=1)                                                   # [N,S]
    vmax = np.take_along_axis(v, imax[:, None, :], 1)[:, 0]
    c = np.zeros(N, dtype)
    for s in range(S):
        c = (c + vmax[:, s]).astype(dtype)
    asel = np.take_along_axis(active, imax[:, None, :], 1)[:, 0]
    gx = np.zeros((N, S, 3), dtype)
    # sign(d)*|d| has zero gradient exactly at d == 0 (torch.sign backward is 0, abs'(0) = 0)
    dsel = np.take_along_axis(d, imax[:, None, :], 1)[:, 0]
    nz = (dsel != 0).astype(dtype)
    for a in range(3):
        gx[..., a] = np.where(imax == a, -1.0, 0.0) + np.where(imax == a + 3, 1.0, 0.0)
    gx = (gx * (asel * nz)[..., None]).astype(dtype)
    return dict(cost=c, gx=gx, occ=occ, imax=imax.astype(np.int32))


def self_term(cost, pw, dtype=np.float32, margin_override=None):
    """Self-collision term on body positions pw [N,L,3] (distance_fields.py:194-215, :113-124)."""
    N = pw.shape[0]
    P = cost.n_pairs
    c = np.zeros(N, dtype)
    gp = np.zeros_like(pw)
    occ = np.zeros(N, bool)
    for p in range(P):
        i, j = int(cost.pair_i[p]), int(cost.pair_j[p])
        v = (pw[:, i] - pw[:, j]).astype(dtype)
        dist = norm3(v[:, 0], v[:, 1], v[:, 2], dtype)
        m = dtype(cost.pair_margin[p])
        occ |= dist < (m if margin_override is None else dtype(margin_override))
        val = (-(dist - m)).astype(dtype)
        act = np.ones(N, dtype)
        if cost.clamp_self:
            act = (val > 0).astype(dtype)
            val = np.maximum(val, 0)
        c = (c + val).astype(dtype)
        with np.errstate(invalid="ignore", divide="ignore"):
            u = np.where(dist[:, None] > 0, v / dist[:, None], 0).astype(dtype)
        gp[:, i] -= u * act[:, None]
        gp[:, j] += u * act[:, None]
    return dict(cost=c, gp=gp, occ=occ)


def ee_term(cost, R, p, H_target=None, dtype=np.float32):
    """EE SE(3) term (geometrics/utils.py:130-154, :209-224; distance_fields.py:347-356).
    R [N,3,3], p [N,3]; H_target [4,4] or [N,4,4].  Returns cost[N], gR[N,3,3], gp[N,3]."""
    Ht = np.asarray(cost.ee_target if H_target is None else H_target, dtype)
    if Ht.ndim == 2:
        Ht = Ht[None]
    Rt, pt = Ht[:, :3, :3], Ht[:, :3, 3]
    w_rot, w_pos = dtype(cost.ee_w_rot), dtype(cost.ee_w_pos)
    N = R.shape[0]
    D = np.zeros(N, dtype)
    gR = np.zeros_like(R)
    gp = np.zeros_like(p)
    if cost.ee_w_rot > 0:
        if dtype == np.float32:
            diag = [fma32(R[:, i, 2], Rt[:, i, 2], fma32(R[:, i, 1], Rt[:, i, 1], R[:, i, 0] * Rt[:, i, 0])) for i in range(3)]
        else:
            diag = [np.sum(R[:, i] * Rt[:, i], -1) for i in range(3)]
        tr = ((diag[0] + diag[1]).astype(dtype) + diag[2]).astype(dtype)
        phi = ((tr - dtype(1.0)) * dtype(0.5)).astype(dtype)
        D = (D + w_rot * (dtype(1.0) - phi)).astype(dtype)
        gR = gR + (-w_rot * dtype(0.5)) * np.broadcast_to(Rt, R.shape)
    dist = None
    if cost.ee_w_pos > 0:
        v = (p - pt).astype(dtype)
        dist = norm3(v[:, 0], v[:, 1], v[:, 2], dtype)
        D = (D + w_pos * dist).astype(dtype)
        with np.errstate(invalid="ignore", divide="ignore"):
            gp = gp + w_pos * np.where(dist[:, None] > 0, v / dist[:, None], 0)
    dist_se3 = D
    if cost.ee_square:
        c = (D * D).astype(dtype)
        scale = dtype(2.0) * D
    else:
        c = D
        scale = np.ones(N, dtype)
    return dict(cost=c, gR=(gR * scale[:, None, None]).astype(dtype), gp=(gp * scale[:, None]).astype(dtype),
                dist=dist_se3)


def objective(desc, cost, q, H_target=None, dtype=np.float32, margin_override=None, want_grad=True):
    """Fused objective: q [N,D] -> per-term costs [N,4] (self, object, ws, ee), weighted total [N],
    grad_q [N,D] of the total, occupancy bits [N] (1 self | 2 object | 4 ws) and nearest-sphere
    indices.  Composition follows tasks/tasks.py:139-232 (+ the EE field added by the caller)."""
    q = np.asarray(q, dtype)
    Rw, pw = fk_forward(desc, q, dtype)
    N = q.shape[0]
    terms = np.zeros((N, 4), dtype)
    gR = np.zeros_like(Rw)
    gp = np.zeros_like(pw)
    occ = np.zeros(N, np.uint8)
    w = cost.weights.astype(dtype)
    out = {}
    if cost.n_pairs:
        r = self_term(cost, pw, dtype, margin_override)
        terms[:, 0] = r["cost"]
        gp += w[0] * r["gp"]
        occ |= r["occ"].astype(np.uint8)
    if cost.n_points:
        x = pw[:, cost.pt_link]
        if cost.n_objects:
            r = object_term(cost, x, dtype, margin_override)
            terms[:, 1] = r["cost"]
            np.add.at(gp, (slice(None), cost.pt_link), w[1] * r["gx"])
            occ |= (r["occ"].astype(np.uint8) << 1)
            out["argmin"], out["sdf"] = r["argmin"], r["sdf"]
        if cost.ws_enabled:
            r = ws_term(cost, x, dtype, margin_override)
            terms[:, 2] = r["cost"]
            np.add.at(gp, (slice(None), cost.pt_link), w[2] * r["gx"])
            occ |= (r["occ"].astype(np.uint8) << 2)
    if cost.ee_enabled:
        e = int(cost.ee_link)
        r = ee_term(cost, Rw[:, e], pw[:, e], H_target, dtype)
        terms[:, 3] = r["cost"]
        gR[:, e] += w[3] * r["gR"]
        gp[:, e] += w[3] * r["gp"]
        out["ee_dist"] = r["dist"]
    total = (((w[0] * terms[:, 0] + w[1] * terms[:, 1]).astype(dtype) + w[2] * terms[:, 2]).astype(dtype)
             + w[3] * terms[:, 3]).astype(dtype)
    out.update(terms=terms, total=total, occ=occ, Rw=Rw, pw=pw)
    if want_grad:
        out["grad_q"] = tree_backward(desc, q, Rw, pw, gR, gp, dtype)
    return out


# ----------------------------------------------------------------------------------------------
# inverse kinematics (robot_tree.py:303-442)
# ----------------------------------------------------------------------------------------------
def ik_loss(desc, ee_link, q, H_target, lower, upper, w_se3=1.0, w_joint_limits=300.0, dtype=np.float32):
    """loss_fn_ik_per_q (robot_tree.py:386-417) and its gradient; also the raw SE(3) error."""
    class _C:
        ee_w_rot, ee_w_pos, ee_square, ee_target = 1.0, 1.0, 0, None
    q = np.asarray(q, dtype)
    Rw, pw = fk_forward(desc, q, dtype)
    r = ee_term(_C, Rw[:, ee_link], pw[:, ee_link], H_target, dtype)
    lo, hi = np.asarray(lower, dtype), np.asarray(upper, dtype)
    below, above = (q < lo), (q > hi)
    e_lim = ((lo - q) ** 2 * below).sum(-1) + ((hi - q) ** 2 * above).sum(-1)
    loss = (dtype(w_se3) * r["cost"] + dtype(w_joint_limits) * e_lim).astype(dtype)
    gR = np.zeros_like(Rw)
    gp = np.zeros_like(pw)
    gR[:, ee_link], gp[:, ee_link] = dtype(w_se3) * r["gR"], dtype(w_se3) * r["gp"]
    grad = tree_backward(desc, q, Rw, pw, gR, gp, dtype)
    grad = grad + dtype(w_joint_limits) * (-2 * (lo - q) * below - 2 * (hi - q) * above)
    return loss, grad.astype(dtype), r["cost"]


def ik_valid(desc, ee_link, q, H_target, lower, upper, se3_eps, dtype=np.float32):
    """ik_termination (robot_tree.py:419-442): sorted indices of seeds inside the limits whose
    SE(3) error is below se3_eps."""
    _, _, err = ik_loss(desc, ee_link, q, H_target, lower, upper, dtype=dtype)
    q = np.asarray(q, dtype)
    ok = np.all((q >= np.asarray(lower, dtype)) & (q <= np.asarray(upper, dtype)), -1) & (err < dtype(se3_eps))
    return np.nonzero(ok)[0].astype(np.int64)


def ik_adam(desc, ee_link, q0, H_target, lower, upper, max_iters, lr, se3_eps, dtype=np.float32):
    """inverse_kinematics (robot_tree.py:303-384) with explicit seeds: torch.optim.Adam defaults
    (betas .9/.999, eps 1e-8, no weight decay), every seed updated every iteration, early exit when
    all seeds are valid (:356)."""
    q = np.asarray(q0, dtype).copy()
    m = np.zeros_like(q)
    v = np.zeros_like(q)
    b1, b2, eps = 0.9, 0.999, 1e-8
    idx = ik_valid(desc, ee_link, q, H_target, lower, upper, se3_eps, dtype)
    for it in range(max_iters):
        idx = ik_valid(desc, ee_link, q, H_target, lower, upper, se3_eps, dtype)
        if idx.size == q.shape[0]:
            break
        _, g, _ = ik_loss(desc, ee_link, q, H_target, lower, upper, dtype=dtype)
        t = it + 1
        m = (b1 * m + (1 - b1) * g).astype(dtype)
        v = (b2 * v + (1 - b2) * g * g).astype(dtype)
        step = lr / (1 - b1 ** t)
        denom = (np.sqrt(v) / np.sqrt(1 - b2 ** t)).astype(dtype) + dtype(eps)
        q = (q - dtype(step) * (m / denom)).astype(dtype)
    return q, idx
