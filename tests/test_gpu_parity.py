"""CUDA kernels (through the C ABI) against the oracle and the golden vectors.  Needs a GPU."""
import numpy as np
import pytest
import torch

from conftest import assert_differences_are_near_ties, ROBOT_NAMES, get_desc, load_golden, planning_cost_from_golden, rel_err, stretch_cost_from_golden
from oracle import restatement as orc

pytestmark = pytest.mark.gpu

POSE_TOL, GRAD_TOL = 1e-5, 1e-4
DEV = "cuda:0"


@pytest.fixture(scope="module")
def models():
    from torch_robotics_b200._lib import DeviceModel
    return {r: DeviceModel(get_desc(r), DEV) for r in ROBOT_NAMES}


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV)


@pytest.mark.parametrize("robot", ROBOT_NAMES)
def test_fk_forward_backward_vs_golden(models, robot):
    g, d, m = load_golden(f"fk_{robot}"), get_desc(robot), models[robot]
    sel = [d.link_names.index(n) for n in g["out_link_names"]]
    sel_arg = None if sel == list(range(d.n_links)) else sel
    H = m.fk_forward(t(g["q"]), sel_arg).cpu().numpy()
    assert H.shape == g["H"].shape
    assert rel_err(H, g["H"]) < POSE_TOL
    assert np.array_equal(H[..., 3, :], np.broadcast_to(np.array([0, 0, 0, 1], np.float32), H[..., 3, :].shape))
    grad = m.fk_backward(t(g["q"]), t(g["G"]), sel_arg).cpu().numpy()
    assert rel_err(grad, g["grad_q"]) < GRAD_TOL


@pytest.mark.parametrize("robot", ["panda", "stretch", "shadow_hand", "tiago_dual_holo_move"])
@pytest.mark.parametrize("n", [1, 31, 33, 1000, 20011])
def test_fk_vs_oracle_ragged_sizes(models, robot, n):
    d, m = get_desc(robot), models[robot]
    rng = np.random.default_rng(n)
    q = (rng.uniform(-3.5, 3.5, (n, d.n_dofs))).astype(np.float32)
    Rw, pw = orc.fk_forward(d, q, np.float32)
    H = m.fk_forward(t(q)).cpu().numpy()
    assert rel_err(H, orc.homogeneous(Rw, pw)) < POSE_TOL
    # a permuted sub-selection of links
    sel = [int(v) for v in rng.permutation(d.n_links)[: max(1, d.n_links // 2)]]
    Hs = m.fk_forward(t(q), sel).cpu().numpy()
    assert np.array_equal(Hs, H[:, sel])
    G = rng.standard_normal(Hs.shape).astype(np.float32)
    grad = m.fk_backward(t(q), t(G), sel).cpu().numpy()
    assert rel_err(grad, orc.fk_backward(d, q, G, sel, np.float32)) < GRAD_TOL


def test_fk_empty_batch(models):
    m = models["panda"]
    H = m.fk_forward(torch.empty((0, 7), device=DEV))
    assert H.shape == (0, 11, 4, 4)


def test_autograd_function_matches_raw(models):
    from torch_robotics_b200 import ops
    g, m = load_golden("fk_panda"), models["panda"]
    q = t(g["q"]).requires_grad_(True)
    H = ops.fk(m, q)
    (H * t(g["G"])).sum().backward()
    assert rel_err(q.grad.cpu().numpy(), g["grad_q"]) < GRAD_TOL


@pytest.mark.parametrize("case", ["panda_spheres3d", "panda_two_objects", "panda_boxes"])
def test_objective_vs_golden(models, case):
    from torch_robotics_b200._lib import DeviceCost
    g, m = load_golden(f"planning_{case}"), models["panda"]
    cost = planning_cost_from_golden(g)
    dc = DeviceCost(m, cost)
    q = t(g["q"].reshape(-1, 7))
    r = dc.objective(q, want_grad=True, want_terms=True, want_occupancy=True, want_argmin=True)
    terms = r["terms"].cpu().numpy()
    for k, name in enumerate(["self", "obj", "ws", "ee"]):
        assert rel_err(terms[:, k], g[f"cost_{name}"].reshape(-1)) < POSE_TOL, name
        w = np.zeros(4, np.float32)
        w[k] = 1
        dc.set_weights(w)
        rk = dc.objective(q)
        assert rel_err(rk["grad_q"].cpu().numpy(), g[f"grad_{name}"].reshape(-1, 7)) < GRAD_TOL, name
        assert rel_err(rk["total"].cpu().numpy(), g[f"cost_{name}"].reshape(-1)) < POSE_TOL
    dc.set_weights([1, 1, 1, 0])
    rc = dc.objective(q, want_occupancy=True)
    assert rel_err(rc["total"].cpu().numpy(), g["cost_collision"].reshape(-1)) < POSE_TOL
    assert rel_err(rc["grad_q"].cpu().numpy(), g["grad_collision"].reshape(-1, 7)) < GRAD_TOL
    # end-to-end occupancy: equal to the reference (no entry of these fixtures sits within a few ulp
    # of a threshold)
    occ = rc["occupancy"].cpu().numpy()
    assert np.array_equal(occ != 0, g["occ_all_default"].reshape(-1))
    occ0 = dc.objective(q, want_grad=False, want_occupancy=True, margin_override=0.0)["occupancy"].cpu().numpy()
    assert np.array_equal((occ0 & 1) != 0, g["occ_self_zero"].reshape(-1))
    assert np.array_equal((occ0 & 2) != 0, g["occ_obj_zero"].reshape(-1))
    assert np.array_equal((occ0 & 4) != 0, g["occ_ws_zero"].reshape(-1))


@pytest.mark.parametrize("case", ["panda_spheres3d", "panda_two_objects", "panda_boxes"])
def test_cost_stage_bit_exact_on_reference_positions(models, case):
    """Stage-wise exactness: fed the reference's link positions, masks, arg-min
    and the decision-chain costs are bit-identical with the oracle, and the masks with the reference."""
    from torch_robotics_b200._lib import DeviceCost
    g, m = load_golden(f"planning_{case}"), models["panda"]
    cost = planning_cost_from_golden(g, with_ee=False)
    dc = DeviceCost(m, cost)
    pw = np.ascontiguousarray(g["link_pos"].reshape(-1, 11, 3))
    x = pw[:, cost.pt_link]
    for tag, mo in (("default", None), ("zero", 0.0)):
        r = dc.cost_stage(t(pw), want_grad=False, want_occupancy=True, want_argmin=True, margin_override=mo)
        occ = r["occupancy"].cpu().numpy()
        assert np.array_equal((occ & 1) != 0, g[f"occ_self_{tag}"].reshape(-1))
        assert np.array_equal((occ & 2) != 0, g[f"occ_obj_{tag}"].reshape(-1))
        assert np.array_equal((occ & 4) != 0, g[f"occ_ws_{tag}"].reshape(-1))
    ro = orc.object_term(cost, x, np.float32)
    assert np.array_equal(r["argmin"].cpu().numpy(), ro["argmin"])
    terms = r["terms"].cpu().numpy()
    assert np.array_equal(terms[:, 0], orc.self_term(cost, pw, np.float32)["cost"])
    assert np.array_equal(terms[:, 1], ro["cost"])
    assert np.array_equal(terms[:, 2], orc.ws_term(cost, x, np.float32)["cost"])


def test_cost_stage_autograd_chain_matches_fused(models):
    """FK (autograd.Function) -> cost stage (autograd.Function) gives the same gradient as the fused kernel."""
    from torch_robotics_b200 import ops
    from torch_robotics_b200._lib import DeviceCost
    g, m = load_golden("planning_panda_spheres3d"), models["panda"]
    dc = DeviceCost(m, planning_cost_from_golden(g))
    q = t(g["q"].reshape(-1, 7)).requires_grad_(True)
    H = ops.fk(m, q)
    c = ops.cost_stage(dc, H[:, :, :3, 3], H[:, 10])
    c.sum().backward()
    fused = dc.objective(q.detach())
    assert rel_err(c.detach().cpu().numpy(), fused["total"].cpu().numpy()) < POSE_TOL
    assert rel_err(q.grad.cpu().numpy(), fused["grad_q"].cpu().numpy()) < GRAD_TOL
    q2 = t(g["q"].reshape(-1, 7)).requires_grad_(True)
    (ops.objective(dc, q2) * 2.0).sum().backward()
    assert rel_err(q2.grad.cpu().numpy(), 2.0 * fused["grad_q"].cpu().numpy()) < 1e-6


def test_objective_vs_oracle_large_with_near_tie_report(models):
    """Fused kernel vs oracle at 100k configs: costs/gradients within tolerance; masks and arg-min
    equal except a handful of entries that sit on a decision boundary (count printed)."""
    from torch_robotics_b200._lib import DeviceCost
    g, d, m = load_golden("planning_panda_spheres3d"), get_desc("panda"), models["panda"]
    cost = planning_cost_from_golden(g)
    dc = DeviceCost(m, cost)
    rng = np.random.default_rng(5)
    lo, hi = d.lo[d.controlled], d.hi[d.controlled]
    q = (lo + rng.random((100_000, 7)) * (hi - lo)).astype(np.float32)
    r = dc.objective(t(q), want_grad=True, want_terms=True, want_occupancy=True, want_argmin=True)
    o = orc.objective(d, cost, q, dtype=np.float32)
    assert rel_err(r["terms"].cpu().numpy(), o["terms"]) < POSE_TOL
    assert rel_err(r["total"].cpu().numpy(), o["total"]) < POSE_TOL
    gd = np.abs(r["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    assert gd[gd < GRAD_TOL].max() < GRAD_TOL
    am, occ = r["argmin"].cpu().numpy(), r["occupancy"].cpu().numpy()
    n_diff = assert_differences_are_near_ties(d, cost, q, am, o["argmin"], occ, o["occ"], np.flatnonzero(gd >= GRAD_TOL), "objective 100k")
    assert n_diff <= 20


def test_stretch_objective(models):
    from torch_robotics_b200._lib import DeviceCost
    g, m = load_golden("objective_stretch"), models["stretch"]
    dc = DeviceCost(m, stretch_cost_from_golden(g))
    r = dc.objective(t(g["q"]), want_terms=True)
    assert rel_err(r["total"].cpu().numpy(), g["cost_total"]) < POSE_TOL
    assert rel_err(r["grad_q"].cpu().numpy(), g["grad_total"]) < GRAD_TOL


def test_rejects_cpu_tensors_and_wrong_dtype(models):
    m = models["panda"]
    with pytest.raises(RuntimeError):
        m.fk_forward(torch.zeros(4, 7))
    with pytest.raises(RuntimeError):
        m.fk_forward(torch.zeros(4, 7, device=DEV, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        m.fk_forward(torch.zeros(4, 6, device=DEV))


def test_decision_sqrt_is_correctly_rounded():
    """sqrt_rn_core (branch-free) == sqrt.rn.f32 bit for bit on 0, on 2^-100..2^126 and on the range
    squared distances actually take."""
    from torch_robotics_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(0)
    n = 1 << 24
    expo = torch.randint(-100, 126, (n,), device=DEV, generator=g).float()
    mant = 1.0 + torch.rand(n, device=DEV, generator=g)
    wide = torch.ldexp(mant, expo.to(torch.int32))
    near = torch.rand(n, device=DEV, generator=g) * 16.0            # squared distances of a few metres
    tiny = torch.rand(n, device=DEV, generator=g) * 1e-6
    special = torch.tensor([0.0, 2.0 ** -100, 1.0, 4.0, 2.0, 3.0, 2.0 ** 125], device=DEV)
    x = torch.cat([wide, near, tiny, special]).contiguous()
    a, b = torch.empty_like(x), torch.empty_like(x)
    _lib.check(_lib.lib().trk_selftest_sqrt(x.data_ptr(), x.numel(), a.data_ptr(), b.data_ptr(), _lib.stream_ptr(x.device)),
               "trk_selftest_sqrt")
    assert torch.equal(a, b)
    assert torch.equal(b, torch.sqrt(x))


def test_posed_nonuniform_objects_bit_exact_vs_oracle(models):
    """Objects with unequal radii / a pose take the MUFU-key scan; value, arg-min and masks must still
    equal the all-exact oracle on identical positions, including deliberately tied spheres."""
    from torch_robotics_b200._lib import DeviceCost
    from torch_robotics_b200.descriptor import build_cost_descriptor
    m = models["panda"]
    rng = np.random.default_rng(11)
    c_a = rng.uniform(-0.8, 0.8, (37, 3)).astype(np.float32)
    r_a = rng.uniform(0.02, 0.3, 37).astype(np.float32)
    c_b = rng.uniform(-0.8, 0.8, (16, 3)).astype(np.float32)
    c_b[5] = c_b[2]                                   # exact duplicate sphere: first index must win
    r_b = np.full(16, 0.1, np.float32)
    c_c = np.array([[0.1, 0.2, 0.3]], np.float32)     # single-sphere object
    cost = build_cost_descriptor(
        pt_link=[2, 3, 5, 7, 9], pt_radius=[0.125, 0.125, 0.13, 0.1, 0.08], cutoff_obj=0.01,
        objects=[(c_a, r_a, None, None), (c_b, r_b, [0.1, -0.2, 0.3], [0.9, 0.1, -0.3, 0.2]), (c_c, [0.25], None, None)],
        ws_min=[-1, -1, -1], ws_max=[1, 1, 1], weights=(0, 1, 1, 0))
    dc = DeviceCost(m, cost)
    n = 200_000
    pw = rng.uniform(-1.0, 1.0, (n, 11, 3)).astype(np.float32)
    pw[:1000, 2] = c_a[3]                             # points exactly on sphere centres
    pw[1000:2000, 3] = c_c[0]
    r = dc.cost_stage(t(pw), want_grad=True, want_occupancy=True, want_argmin=True)
    x = pw[:, cost.pt_link]
    ro = orc.object_term(cost, x, np.float32)
    rw = orc.ws_term(cost, x, np.float32)
    assert np.array_equal(r["argmin"].cpu().numpy(), ro["argmin"])
    terms = r["terms"].cpu().numpy()
    assert np.array_equal(terms[:, 1], ro["cost"])
    assert np.array_equal(terms[:, 2], rw["cost"])
    occ = r["occupancy"].cpu().numpy()
    assert np.array_equal((occ & 2) != 0, ro["occ"]) and np.array_equal((occ & 4) != 0, rw["occ"])
    gp = np.zeros_like(pw)
    np.add.at(gp, (slice(None), cost.pt_link), ro["gx"] + rw["gx"])
    assert rel_err(r["grad_pos"].cpu().numpy(), gp) < 1e-5


def test_box_obstacles_bit_exact_vs_oracle_at_scale(models):
    """Rounded and sharp boxes mixed with spheres in posed objects: decision chain bit-identical with the oracle
    on 200k random position sets, including points inside boxes and exactly on faces / centres."""
    from torch_robotics_b200._lib import DeviceCost
    g, m = load_golden("planning_panda_boxes"), models["panda"]
    cost = planning_cost_from_golden(g, with_ee=False)
    cost.weights = np.array([0, 1, 0, 0], np.float32)
    dc = DeviceCost(m, cost)
    rng = np.random.default_rng(21)
    n = 200_000
    pw = rng.uniform(-0.9, 0.9, (n, 11, 3)).astype(np.float32)
    pw[:500, 2] = cost.box_center[0]                                  # box centre
    pw[500:1000, 3] = cost.box_center[1] + cost.box_half[1] * np.array([1, 0, 0], np.float32)   # on a face
    r = dc.cost_stage(t(pw), want_grad=True, want_occupancy=True, want_argmin=True)
    x = pw[:, cost.pt_link]
    ro = orc.object_term(cost, x, np.float32)
    assert np.array_equal(r["argmin"].cpu().numpy(), ro["argmin"])
    assert np.array_equal(r["terms"].cpu().numpy()[:, 1], ro["cost"])
    assert np.array_equal((r["occupancy"].cpu().numpy() & 2) != 0, ro["occ"])
    gp = np.zeros_like(pw)
    np.add.at(gp, (slice(None), cost.pt_link), ro["gx"])
    assert rel_err(r["grad_pos"].cpu().numpy(), gp) < 1e-5


def test_new_entry_points_validate_arguments(models):
    """Argument errors of the later-added entry points surface as RuntimeError with the library's message."""
    import ctypes
    from torch_robotics_b200 import _lib
    from torch_robotics_b200._lib import DeviceCost
    from torch_robotics_b200.descriptor import VoxelSDF, build_cost_descriptor
    m = models["panda"]
    L = _lib.lib()
    q = torch.zeros(4, 7, device=DEV)
    J = torch.empty(4 * 11 * 7 * 7 + 1, device=DEV)
    assert L.trk_link_jacobian(m.handle, q.data_ptr(), 4, J.data_ptr() + 4, None) == -3            # misaligned output
    assert b"16-byte" in L.trk_last_error()
    assert L.trk_link_jacobian(m.handle, q.data_ptr(), -1, J.data_ptr(), None) == -2
    assert m.link_jacobian(torch.empty(0, 7, device=DEV)).shape == (0, 11, 7, 7)
    with pytest.raises(RuntimeError):
        m.link_jacobian(torch.zeros(4, 6, device=DEV))
    # voxel SDF: the grid object must own no primitives, have the identity pose and a positive extent
    grid = np.zeros((2, 2, 2, 4), np.float32)
    good = build_cost_descriptor(pt_link=[2], pt_radius=[0.1], objects=[VoxelSDF((2, 2, 2), np.zeros(3), np.ones(3), grid)])
    DeviceCost(m, good)
    bad = build_cost_descriptor(pt_link=[2], pt_radius=[0.1], objects=[VoxelSDF((2, 2, 2), np.zeros(3), np.array([1.0, 0.0, 1.0]), grid)])
    with pytest.raises(RuntimeError, match="extent"):
        DeviceCost(m, bad)
    wrong_shape = build_cost_descriptor(pt_link=[2], pt_radius=[0.1], objects=[VoxelSDF((3, 2, 2), np.zeros(3), np.ones(3), grid)])
    with pytest.raises(RuntimeError, match="voxel SDF data"):
        DeviceCost(m, wrong_shape)
    # objective / IK on an empty batch
    r = DeviceCost(m, good).objective(torch.empty(0, 7, device=DEV))
    assert r["total"].shape == (0,) and r["grad_q"].shape == (0, 7)


@pytest.mark.parametrize("robot,n_pts", [("panda", 7), ("ur10", 3), ("shadow_hand", 11), ("stretch", 9)])
def test_table_driven_static_objective_joint_scan_vs_oracle(models, robot, n_pts):
    """Compiled-in robots with collision points / pairs that are NOT a compile-time collision model go through the
    table-driven k_objective_static; with one equal-radius sphere object its points are scanned jointly in run-time
    groups of five (ragged last group).  Costs, gradients, masks and arg-min against the C oracle."""
    from oracle import c_oracle as co
    from torch_robotics_b200._lib import DeviceCost
    from torch_robotics_b200.descriptor import build_cost_descriptor
    d, m = get_desc(robot), models[robot]
    assert m.kernel_family != "generic"
    rng = np.random.default_rng(n_pts)
    pts = [int(v) for v in rng.permutation(np.arange(1, d.n_links))[:n_pts]]
    pairs = [(pts[0], 0), (pts[-1], pts[1])] if n_pts > 2 else []
    centers = rng.uniform(-0.6, 0.6, (9, 3)).astype(np.float32)
    cost = build_cost_descriptor(pt_link=pts, pt_radius=rng.uniform(0.03, 0.1, n_pts).astype(np.float32), cutoff_obj=0.02,
                                 objects=[(centers, np.full(9, 0.1, np.float32), None, None)], ws_min=[-0.8, -0.8, -0.2], ws_max=[0.8, 0.8, 1.2],
                                 self_pairs=pairs, self_margin=[0.04], ee_link=d.n_links - 1, ee_target=np.eye(4, dtype=np.float32))
    lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float32)
    hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float32)
    q = (lo + rng.random((30_011, d.n_dofs), dtype=np.float32) * (hi - lo)).astype(np.float32)
    r = DeviceCost(m, cost).objective(t(q), want_terms=True, want_occupancy=True, want_argmin=True)
    o = co.objective(co.Tree(d), co.Cost(cost), q)
    assert rel_err(r["terms"].cpu().numpy(), o["terms"]) < POSE_TOL
    gd = np.abs(r["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    assert gd[gd < GRAD_TOL].max() < GRAD_TOL
    n_diff = assert_differences_are_near_ties(d, cost, q, r["argmin"].cpu().numpy(), o["argmin"], r["occupancy"].cpu().numpy(), o["occ"],
                                              np.flatnonzero(gd >= GRAD_TOL), f"table-driven {robot}")
    assert n_diff <= 30
    fwd = DeviceCost(m, cost).objective(t(q), want_grad=False, want_terms=False, want_occupancy=True, margin_override=0.0)
    o0 = co.objective(co.Tree(d), co.Cost(cost), q, margin_override=0.0)
    zero = (np.zeros_like(cost.pair_margin), np.zeros_like(cost.pt_margin_obj), np.zeros_like(cost.pt_margin_ws))
    assert_differences_are_near_ties(d, cost, q, None, None, fwd["occupancy"].cpu().numpy(), o0["occ"], [], f"table-driven {robot}, margin 0",
                                     occ_margins=zero)
    assert torch.equal(fwd["total"], r["total"])
