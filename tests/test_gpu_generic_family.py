"""The runtime-tree ("generic") kernel family - what serves any URDF outside the compiled-in set and every tree with
attached points - exercised on all shipped robots by switching the compiled-in families off
(TRK_B200_FORCE_GENERIC, read by trk_model_create)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROBOT_NAMES, get_desc, load_golden, planning_cost_from_golden, rel_err, stretch_cost_from_golden
from oracle import c_oracle as co
from oracle import restatement as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def generic_models():
    from torch_robotics_b200._lib import DeviceModel
    before = os.environ.get("TRK_B200_FORCE_GENERIC")
    os.environ["TRK_B200_FORCE_GENERIC"] = "1"
    try:
        models = {r: DeviceModel(get_desc(r), DEV) for r in ROBOT_NAMES}
    finally:
        if before is None:
            del os.environ["TRK_B200_FORCE_GENERIC"]
        else:
            os.environ["TRK_B200_FORCE_GENERIC"] = before      # the whole suite may be run with the switch set
    for m in models.values():
        assert m.kernel_family == "generic"
    return models


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV)


@pytest.mark.parametrize("robot", ROBOT_NAMES)
def test_fk_forward_backward_vs_golden(generic_models, robot):
    g, d, m = load_golden(f"fk_{robot}"), get_desc(robot), generic_models[robot]
    sel = [d.link_names.index(n) for n in g["out_link_names"]]
    sel_arg = None if sel == list(range(d.n_links)) else sel
    H = m.fk_forward(t(g["q"]), sel_arg).cpu().numpy()
    assert rel_err(H, g["H"]) < 1e-5
    grad = m.fk_backward(t(g["q"]), t(g["G"]), sel_arg).cpu().numpy()
    assert rel_err(grad, g["grad_q"]) < 1e-4


@pytest.mark.parametrize("robot", ROBOT_NAMES)
@pytest.mark.parametrize("n", [1, 33, 4099])
def test_fk_vs_oracle_ragged_and_subselection(generic_models, robot, n):
    d, m = get_desc(robot), generic_models[robot]
    rng = np.random.default_rng(n)
    q = rng.uniform(-3.5, 3.5, (n, d.n_dofs)).astype(np.float32)
    H = m.fk_forward(t(q)).cpu().numpy()
    assert rel_err(H, co.fk_forward(co.Tree(d), q)) < 1e-5
    sel = [int(v) for v in rng.permutation(d.n_links)[: max(1, d.n_links // 2)]]
    Hs = m.fk_forward(t(q), sel).cpu().numpy()
    assert np.array_equal(Hs, H[:, sel])
    G = rng.standard_normal(Hs.shape).astype(np.float32)
    grad = m.fk_backward(t(q), t(G), sel).cpu().numpy()
    assert rel_err(grad, co.fk_backward(co.Tree(d), q, G, sel)) < 1e-4
    # clamped joints: exactly zero gradient
    _, mask = orc.joint_angles(d, q, np.float32)
    for b in range(d.n_links):
        if d.q_index[b] >= 0:
            assert np.all(grad[mask[:, b] == 0, d.q_index[b]] == 0)


@pytest.mark.parametrize("case", ["panda_spheres3d", "panda_two_objects", "panda_boxes"])
def test_objective_vs_golden_and_static_family(generic_models, case):
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    g, d = load_golden(f"planning_{case}"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    q = t(g["q"].reshape(-1, 7))
    r = DeviceCost(generic_models["panda"], cost).objective(q, want_terms=True, want_occupancy=True, want_argmin=True)
    terms = r["terms"].cpu().numpy()
    for k, name in enumerate(("self", "obj", "ws", "ee")):
        assert rel_err(terms[:, k], g[f"cost_{name}"].reshape(-1)) < 1e-5
    # decisions are shared with the compiled-in family bit for bit (same positions would be needed for the costs)
    s = DeviceCost(DeviceModel(d, DEV), cost).objective(q, want_terms=True, want_occupancy=True, want_argmin=True)
    assert np.mean(r["occupancy"].cpu().numpy() != s["occupancy"].cpu().numpy()) < 0.01
    assert np.mean(r["argmin"].cpu().numpy() != s["argmin"].cpu().numpy()) < 0.01
    assert rel_err(r["total"].cpu().numpy(), s["total"].cpu().numpy()) < 1e-5
    assert rel_err(r["grad_q"].cpu().numpy(), s["grad_q"].cpu().numpy()) < 1e-4
    fwd = DeviceCost(generic_models["panda"], cost).objective(q, want_grad=False, want_terms=False)
    assert torch.equal(fwd["total"], r["total"]) and fwd["grad_q"] is None


def test_objective_vs_oracle_at_scale(generic_models):
    from torch_robotics_b200._lib import DeviceCost
    for robot, cost in (("panda", planning_cost_from_golden(load_golden("planning_panda_spheres3d"))),
                        ("stretch", stretch_cost_from_golden(load_golden("objective_stretch")))):
        d, m = get_desc(robot), generic_models[robot]
        rng = np.random.default_rng(3)
        lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float32)
        hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float32)
        q = (lo + rng.random((100_003, d.n_dofs), dtype=np.float32) * (hi - lo) * 1.1 - 0.05 * (hi - lo)).astype(np.float32)
        r = DeviceCost(m, cost).objective(t(q), want_terms=True, want_occupancy=True, want_argmin=True)
        o = co.objective(co.Tree(d), co.Cost(cost), q)
        assert rel_err(r["total"].cpu().numpy(), o["total"]) < 1e-5
        gd = np.abs(r["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
        assert (gd >= 1e-4).sum() <= 30 and gd[gd < 1e-4].max() < 1e-4
        assert np.mean(r["occupancy"].cpu().numpy() != o["occ"]) < 1e-4
        assert np.mean(r["argmin"].cpu().numpy() != o["argmin"]) < 1e-4


def test_ee_target_per_configuration_and_margin_override(generic_models):
    from torch_robotics_b200._lib import DeviceCost
    g, d, m = load_golden("planning_panda_spheres3d"), get_desc("panda"), generic_models["panda"]
    cost = planning_cost_from_golden(g)
    q = g["q"].reshape(-1, 7)
    rng = np.random.default_rng(9)
    tgt = np.tile(g["H_target"].reshape(1, 4, 4), (q.shape[0], 1, 1)).astype(np.float32)
    tgt[:, :3, 3] += rng.normal(0, 0.2, (q.shape[0], 3)).astype(np.float32)
    r = DeviceCost(m, cost).objective(t(q), ee_target_cfg=t(tgt), want_occupancy=True, margin_override=0.0)
    o = co.objective(co.Tree(d), co.Cost(cost), q, ee_target_cfg=tgt, margin_override=0.0)
    assert rel_err(r["total"].cpu().numpy(), o["total"]) < 1e-5
    assert rel_err(r["grad_q"].cpu().numpy(), o["grad_q"]) < 1e-4
    assert np.array_equal(r["occupancy"].cpu().numpy(), o["occ"])


def test_fused_ik_matches_golden_and_compiled_in_family(generic_models):
    """k_ik_adam_generic: iterates equal the reference's Adam loop (golden ik_panda.npz) and the register-resident
    kernel's; termination masks and per-iteration valid counts are identical."""
    from torch_robotics_b200._lib import DeviceModel
    g, d = load_golden("ik_panda"), get_desc("panda")
    mg, ms = generic_models["panda"], DeviceModel(d, DEV)
    q0 = t(g["q0"])
    for iters in (1, 5, 40):
        kw = dict(n_iters=iters + 1, n_updates=iters, lr=0.2, se3_eps=0.05, want_loss_stats=True)
        rg = mg.ik_adam(q0, g["H_target"], 10, g["lower"], g["upper"], **kw)
        rs = ms.ik_adam(q0, g["H_target"], 10, g["lower"], g["upper"], **kw)
        assert rel_err(rg["q"].cpu().numpy(), g[f"q_ik_{iters}"]) < (1e-4 if iters < 40 else 5e-3)
        assert rel_err(rg["q"].cpu().numpy(), rs["q"].cpu().numpy()) < (1e-4 if iters < 40 else 5e-3)
        if iters < 40:
            assert torch.equal(rg["valid"], rs["valid"]) and torch.equal(rg["valid_count"], rs["valid_count"])
            assert np.array_equal(torch.nonzero(rg["valid"]).reshape(-1).cpu().numpy(), g[f"idx_valid_{iters}"])
        assert rel_err(rg["loss_sum"].cpu().numpy(), rs["loss_sum"].cpu().numpy()) < 1e-3


def test_fused_ik_converges_on_a_tree_with_prismatic_joints(generic_models):
    """Stretch (prismatic lift / telescope, branches): reachable targets are reached from perturbed seeds."""
    d, m = get_desc("stretch"), generic_models["stretch"]
    ee = d.link_names.index("link_straight_gripper")
    rng = np.random.default_rng(21)
    lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float32)
    hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float32)
    q_goal = (lo + rng.random((64, d.n_dofs), dtype=np.float32) * (hi - lo) * 0.6 + 0.2 * (hi - lo)).astype(np.float32)
    Rw, pw = orc.fk_forward(d, q_goal, np.float32)
    H = orc.homogeneous(Rw, pw, [ee])[:, 0]
    q0 = np.clip(q_goal + rng.normal(0, 0.1, q_goal.shape).astype(np.float32), lo, hi)
    r = m.ik_adam(t(q0), t(H), ee, lo, hi, n_iters=300, lr=0.02, se3_eps=0.02)
    assert r["valid"].float().mean().item() > 0.9
    assert int(r["valid_count"][-1]) == int(r["valid"].sum())


def test_base_pose_is_a_rigid_motion_of_the_scene():
    """A robot with a base pose B (served by the runtime-tree kernels: the compiled-in tables have the identity root)
    sees obstacles O and an EE target T exactly as the identity-base robot sees B^-1 O and B^-1 T: same costs, same
    gradients; link poses and Jacobians are those of the identity-base robot moved by B."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    from torch_robotics_b200.descriptor import build_cost_descriptor, quat_wxyz_to_matrix
    import copy
    d0 = get_desc("panda")
    quat = np.array([0.8, 0.2, -0.4, 0.4], np.float32)
    quat /= np.linalg.norm(quat)
    Rb, tb = quat_wxyz_to_matrix(quat).astype(np.float64), np.array([0.3, -0.2, 0.15])
    d1 = copy.deepcopy(d0)
    d1.base_R, d1.base_t = Rb.astype(np.float32), tb.astype(np.float32)
    m0, m1 = DeviceModel(d0, DEV), DeviceModel(d1, DEV)
    assert m0.kernel_family == "panda" and m1.kernel_family == "generic"
    rng = np.random.default_rng(17)
    q = rng.uniform(-2.0, 2.0, (3001, 7)).astype(np.float32)
    H0, H1 = m0.fk_forward(t(q)).cpu().numpy().astype(np.float64), m1.fk_forward(t(q)).cpu().numpy()
    B = np.eye(4)
    B[:3, :3], B[:3, 3] = Rb, tb
    assert rel_err(H1, B @ H0) < 1e-5
    # objective: world-frame spheres / EE target for the moved robot, pulled back by B^-1 for the identity-base robot
    centers_w = rng.uniform(-0.6, 0.6, (7, 3))
    radii = np.full(7, 0.12, np.float32)
    Tw = np.eye(4)
    Tw[:3, :3], Tw[:3, 3] = quat_wxyz_to_matrix(np.array([0.6, -0.3, 0.1, 0.7], np.float32) / np.float32(0.9746794)), [0.4, 0.1, 0.5]
    Binv = np.linalg.inv(B)
    pairs = [(9, 0), (9, 1), (4, 1), (6, 2)]
    common = dict(pt_link=[2, 3, 5, 7, 9], pt_radius=[0.125, 0.125, 0.13, 0.1, 0.08], cutoff_obj=0.01, self_pairs=pairs,
                  self_margin=[0.05], ee_link=10)
    c1 = build_cost_descriptor(objects=[(centers_w.astype(np.float32), radii, None, None)], ee_target=Tw.astype(np.float32), **common)
    c0 = build_cost_descriptor(objects=[((centers_w @ Binv[:3, :3].T + Binv[:3, 3]).astype(np.float32), radii, None, None)],
                               ee_target=(Binv @ Tw).astype(np.float32), **common)
    r1 = DeviceCost(m1, c1).objective(t(q), want_terms=True)
    r0 = DeviceCost(m0, c0).objective(t(q), want_terms=True)
    assert rel_err(r1["terms"].cpu().numpy(), r0["terms"].cpu().numpy()) < 2e-5
    assert rel_err(r1["grad_q"].cpu().numpy(), r0["grad_q"].cpu().numpy()) < 1e-4
    # Jacobians: positions rotate with B
    J0, J1 = m0.link_jacobian(t(q[:200])).cpu().numpy().astype(np.float64), m1.link_jacobian(t(q[:200])).cpu().numpy()
    assert rel_err(J1[:, :, 0:3, :], np.einsum("ij,nljd->nlid", Rb, J0[:, :, 0:3, :])) < 1e-5
