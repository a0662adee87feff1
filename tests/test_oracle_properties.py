"""Size-independent properties of the path, checked on the CPU oracles (no GPU): they hold for any robot and any batch
size, so they complement the golden vectors (which pin the arithmetic) at sizes the goldens do not cover."""
import copy

import numpy as np
import pytest

from conftest import ROBOT_NAMES, get_desc, load_golden, planning_cost_from_golden, rel_err
from oracle import c_oracle as co
from oracle import restatement as orc


def uniform_q(d, n, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float64)
    hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float64)
    mid, half = (lo + hi) / 2, (hi - lo) / 2 * spread
    return (mid + (rng.random((n, d.n_dofs)) * 2 - 1) * half).astype(np.float32)


@pytest.mark.parametrize("robot", ROBOT_NAMES)
def test_rotations_stay_orthonormal_and_rows_are_independent(robot):
    d = get_desc(robot)
    q = uniform_q(d, 257, 1, spread=1.3)                      # includes out-of-limit rows (clamped)
    Rw, pw = orc.fk_forward(d, q, np.float64)
    eye = np.broadcast_to(np.eye(3), Rw.shape)
    # the joint-origin rotations are fp32 constants (rigid_body.py:93 builds them in fp32), so orthonormality holds to
    # fp32 accuracy times the chain length, not to fp64
    assert np.abs(Rw @ np.swapaxes(Rw, -1, -2) - eye).max() < 2e-6
    assert np.abs(np.linalg.det(Rw) - 1).max() < 2e-6
    # batch rows do not interact: any sub-batch gives the same rows
    Rs, ps = orc.fk_forward(d, q[100:120], np.float64)
    assert np.array_equal(Rs, Rw[100:120]) and np.array_equal(ps, pw[100:120])


@pytest.mark.parametrize("robot", ["panda", "stretch", "shadow_hand", "tiago_dual_holo_move"])
def test_reverse_mode_is_the_adjoint_of_the_forward_tangent(robot):
    """<G, dH/dq . v> == <grad_q(G), v> with the tangent taken by central differences in fp64 (inside the limits)."""
    d = get_desc(robot)
    rng = np.random.default_rng(2)
    q = uniform_q(d, 8, 3, spread=0.8).astype(np.float64)
    v = rng.standard_normal(q.shape)
    G = rng.standard_normal((8, d.n_links, 4, 4))
    G[:, :, 3, :] = 0
    eps = 1e-6
    Hp = orc.homogeneous(*orc.fk_forward(d, q + eps * v, np.float64))
    Hm = orc.homogeneous(*orc.fk_forward(d, q - eps * v, np.float64))
    lhs = np.sum(G * (Hp - Hm) / (2 * eps), axis=(1, 2, 3))
    rhs = np.sum(orc.fk_backward(d, q, G, None, np.float64) * v, axis=1)
    assert np.abs(lhs - rhs).max() / np.abs(lhs).max() < 1e-7


def test_objective_is_invariant_under_a_rigid_motion_of_robot_and_scene():
    """Moving the robot base and the whole scene by the same rigid motion changes neither costs nor gradients."""
    from torch_robotics_b200.descriptor import build_cost_descriptor, quat_wxyz_to_matrix
    d0 = get_desc("panda")
    quat = np.array([0.5, -0.5, 0.1, 0.7], np.float32)
    Rb, tb = quat_wxyz_to_matrix(quat / np.linalg.norm(quat)).astype(np.float64), np.array([0.2, 0.4, -0.3])
    rng = np.random.default_rng(4)
    centers = rng.uniform(-0.6, 0.6, (6, 3))
    T = np.eye(4)
    T[:3, 3] = [0.3, 0.2, 0.5]
    common = dict(pt_link=[2, 3, 5, 7, 9], pt_radius=[0.1] * 5, cutoff_obj=0.01, self_pairs=[(9, 0), (6, 1)], self_margin=[0.05], ee_link=10)
    c0 = build_cost_descriptor(objects=[(centers.astype(np.float32), np.full(6, 0.1, np.float32), None, None)], ee_target=T.astype(np.float32), **common)
    B = np.eye(4)
    B[:3, :3], B[:3, 3] = Rb, tb
    c1 = build_cost_descriptor(objects=[((centers @ Rb.T + tb).astype(np.float32), np.full(6, 0.1, np.float32), None, None)],
                               ee_target=(B @ T).astype(np.float32), **common)
    q = uniform_q(d0, 300, 5)
    o0 = orc.objective(d0, c0, q, dtype=np.float64)
    # the moved robot: same tree, base pose B (the restatement takes the base as arguments of fk_forward)
    Rw, pw = orc.fk_forward(d0, q, np.float64, base_R=Rb, base_t=tb)
    Rw0, pw0 = orc.fk_forward(d0, q, np.float64)
    assert np.abs(pw - (pw0 @ Rb.T + tb)).max() < 1e-12
    x1 = pw[:, c1.pt_link]
    x0 = pw0[:, c0.pt_link]
    t1, t0 = orc.object_term(c1, x1, np.float64), orc.object_term(c0, x0, np.float64)
    assert rel_err(t1["cost"], t0["cost"]) < 1e-6 and np.array_equal(t1["argmin"], t0["argmin"])
    assert rel_err(t1["gx"], t0["gx"] @ Rb.T) < 1e-6          # forces rotate with the scene
    assert np.isfinite(o0["total"]).all()


def test_attached_point_equals_manual_transform_and_receives_the_right_wrench():
    """A point attached with offset o to link l is R_w[l] o + p_w[l]; a force on it back-propagates like the same force
    plus the torque o_w x f on l (here: through the position gradient of an equivalent upstream G)."""
    from torch_robotics_b200.descriptor import attach_points
    d = get_desc("panda")
    off = np.array([0.07, -0.02, 0.11], np.float32).astype(np.float64)      # offsets are fp32 constants of the tree
    d2, ids = attach_points(d, [("panda_link5", off)])
    q = uniform_q(d, 64, 6).astype(np.float64)
    Rw, pw = orc.fk_forward(d2, q, np.float64)
    l = d.link_names.index("panda_link5")
    assert np.abs(pw[:, ids[0]] - (np.einsum("nij,j->ni", Rw[:, l], off) + pw[:, l])).max() < 1e-12
    rng = np.random.default_rng(7)
    f = rng.standard_normal((64, 3))
    G2 = np.zeros((64, d2.n_links, 4, 4))
    G2[:, ids[0], :3, 3] = f
    g_point = orc.fk_backward(d2, q, G2, None, np.float64)
    # the same gradient expressed on link l: d(R o + p)/dq . f  ->  G_R = f o^T, g_p = f
    G1 = np.zeros((64, d.n_links, 4, 4))
    G1[:, l, :3, 3] = f
    G1[:, l, :3, :3] = f[:, :, None] * off[None, None, :]
    g_link = orc.fk_backward(d, q, G1, None, np.float64)
    assert np.abs(g_point - g_link).max() < 1e-12


def test_c_oracle_and_numpy_restatement_agree_at_a_size_the_goldens_do_not_cover():
    g, d = load_golden("planning_panda_two_objects"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    q = uniform_q(d, 20_000, 8)
    a = orc.objective(d, cost, q, dtype=np.float32)
    b = co.objective(co.Tree(d), co.Cost(cost), q)
    assert rel_err(a["total"], b["total"]) < 1e-6 and rel_err(a["grad_q"], b["grad_q"]) < 1e-5
    assert np.array_equal(a["occ"].astype(np.uint8), b["occ"]) and np.array_equal(a["argmin"], b["argmin"])
