"""The numpy restatement (oracle/restatement.py) against golden vectors recorded from the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import ROBOT_NAMES, get_desc, load_golden, planning_cost_from_golden, rel_err, stretch_cost_from_golden
from oracle import restatement as orc

POSE_TOL = 1e-5     # poses / costs 1e-5 relative
GRAD_TOL = 1e-4     # gradients 1e-4 relative


@pytest.mark.parametrize("robot", ROBOT_NAMES)
def test_descriptor_matches_reference_bodies(robot):
    g, d = load_golden(f"fk_{robot}"), get_desc(robot)
    assert list(g["link_names"]) == d.link_names
    assert np.array_equal(g["parent"], d.parent)
    assert np.array_equal(g["fixed_rotation"], d.R_fixed)      # bit-exact fp32 constants
    assert np.array_equal(g["trans"], d.trans)
    assert np.array_equal(g["joint_axis"], d.axis)
    assert np.array_equal(g["joint_idx"], d.q_index)
    assert list(g["joint_type"]) == d.joint_types
    for l in range(d.n_links):
        if g["has_limit_dict"][l]:
            assert d.limits[l]["lower"] == g["limit_lower"][l] and d.limits[l]["upper"] == g["limit_upper"][l]
        else:
            assert d.limits[l] is None


@pytest.mark.parametrize("robot", ROBOT_NAMES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fk_forward_and_backward(robot, dtype):
    g, d = load_golden(f"fk_{robot}"), get_desc(robot)
    names = d.link_names
    sel = [names.index(n) for n in g["out_link_names"]]
    Rw, pw = orc.fk_forward(d, g["q"], dtype)
    H = orc.homogeneous(Rw, pw, sel)
    assert H.shape == g["H"].shape
    assert rel_err(H, g["H"]) < POSE_TOL
    grad = orc.fk_backward(d, g["q"], g["G"], sel, dtype)
    assert rel_err(grad, g["grad_q"]) < GRAD_TOL


@pytest.mark.parametrize("case", ["panda_spheres3d", "panda_two_objects", "panda_boxes"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_planning_terms(case, dtype):
    g, d = load_golden(f"planning_{case}"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    q = g["q"].reshape(-1, 7)
    r = orc.objective(d, cost, q, dtype=dtype)
    assert rel_err(r["pw"], g["link_pos"].reshape(-1, 11, 3)) < POSE_TOL
    for t, name in enumerate(["self", "obj", "ws", "ee"]):
        assert rel_err(r["terms"][:, t], g[f"cost_{name}"].reshape(-1)) < POSE_TOL, name
        w = np.zeros(4, np.float32)
        w[t] = 1
        cost.weights = w
        rt = orc.objective(d, cost, q, dtype=dtype)
        assert rel_err(rt["grad_q"], g[f"grad_{name}"].reshape(-1, 7)) < GRAD_TOL, name
    cost.weights = np.array([1, 1, 1, 0], np.float32)
    rc = orc.objective(d, cost, q, dtype=dtype)
    assert rel_err(rc["total"], g["cost_collision"].reshape(-1)) < POSE_TOL
    assert rel_err(rc["grad_q"], g["grad_collision"].reshape(-1, 7)) < GRAD_TOL


@pytest.mark.parametrize("case", ["panda_spheres3d", "panda_two_objects", "panda_boxes"])
def test_planning_masks_and_argmin_bit_exact(case):
    """Stage-wise exactness: fed the reference's own link positions, the cost stage reproduces the
    reference's occupancy masks and nearest-sphere indices exactly."""
    g = load_golden(f"planning_{case}")
    cost = planning_cost_from_golden(g)
    pw = g["link_pos"].reshape(-1, 11, 3)
    x = pw[:, cost.pt_link]
    for tag, mo in (("default", None), ("zero", 0.0)):
        assert np.array_equal(orc.self_term(cost, pw, np.float32, mo)["occ"], g[f"occ_self_{tag}"].reshape(-1))
        ro = orc.object_term(cost, x, np.float32, mo)
        assert np.array_equal(ro["occ"], g[f"occ_obj_{tag}"].reshape(-1))
        assert np.array_equal(orc.ws_term(cost, x, np.float32, mo)["occ"], g[f"occ_ws_{tag}"].reshape(-1))
    n_obj = int(g["n_obj"])
    identity = [bool(cost.obj_identity[o]) for o in range(n_obj)]
    am = g["obj_argmin"].reshape(-1, n_obj, cost.n_points)
    for o in range(n_obj):
        if (am[:, o] < 0).all():
            continue            # arg-min fixtures exist for sphere-only objects
        if identity[o]:
            assert np.array_equal(ro["argmin"][:, o], am[:, o])
        else:   # posed object: the reference goes through a BLAS matmul, allow near-tie flips only
            diff = ro["argmin"][:, o] != am[:, o]
            assert diff.mean() < 0.01


def test_stretch_objective():
    g, d = load_golden("objective_stretch"), get_desc("stretch")
    cost = stretch_cost_from_golden(g)
    for dtype in (np.float32, np.float64):
        r = orc.objective(d, cost, g["q"], dtype=dtype)
        assert rel_err(r["terms"][:, 1], g["cost_obj"]) < POSE_TOL
        assert rel_err(r["terms"][:, 2], g["cost_ws"]) < POSE_TOL
        assert rel_err(r["terms"][:, 3], g["cost_ee"]) < POSE_TOL
        assert rel_err(r["total"], g["cost_total"]) < POSE_TOL
        assert rel_err(r["grad_q"], g["grad_total"]) < GRAD_TOL
    pw = orc.fk_forward(d, g["q"], np.float32)[1]
    assert np.array_equal((r["occ"] >> 1) & 1, g["occ_obj"].astype(np.uint8)) or \
        np.mean(((r["occ"] >> 1) & 1) != g["occ_obj"]) < 0.05
    assert pw.shape == (32, 44, 3)


def test_ik_loss_termination_and_adam():
    g, d = load_golden("ik_panda"), get_desc("panda")
    loss, grad, _ = orc.ik_loss(d, 10, g["q_probe"], g["H_target"], g["lower"], g["upper"])
    assert rel_err(loss, g["loss"]) < POSE_TOL
    assert rel_err(grad, g["loss_grad"]) < GRAD_TOL
    idx = orc.ik_valid(d, 10, g["q_probe"], g["H_target"], g["lower"], g["upper"], 0.5)
    assert np.array_equal(idx, g["idx_valid_eps0p5"])
    for iters in (1, 5, 40):
        q, idx = orc.ik_adam(d, 10, g["q0"], g["H_target"], g["lower"], g["upper"], iters, 0.2, 0.05)
        assert rel_err(q, g[f"q_ik_{iters}"]) < (1e-4 if iters < 40 else 5e-3)
        if iters < 40:
            assert np.array_equal(idx, g[f"idx_valid_{iters}"])


def test_fma32_is_exact():
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.fmaf.restype = ctypes.c_float
    libm.fmaf.argtypes = [ctypes.c_float] * 3
    rng = np.random.default_rng(0)
    a = rng.standard_normal(20000).astype(np.float32)
    b = rng.standard_normal(20000).astype(np.float32)
    c = (-(a.astype(np.float64) * b).astype(np.float32) * rng.choice([1.0, 0.5, 1e-3], 20000)).astype(np.float32)
    ref = np.array([libm.fmaf(float(x), float(y), float(z)) for x, y, z in zip(a, b, c)], np.float32)
    assert np.array_equal(orc.fma32(a, b, c), ref)
