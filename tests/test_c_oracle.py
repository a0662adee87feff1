"""The plain-C restatement (oracle/trk_oracle.c): golden vectors + bit-exact agreement with the numpy
restatement on every decision value.  CPU only."""
import numpy as np
import pytest

from conftest import ROBOT_NAMES, get_desc, load_golden, planning_cost_from_golden, rel_err, stretch_cost_from_golden
from oracle import c_oracle as co
from oracle import restatement as orc


@pytest.mark.parametrize("robot", ROBOT_NAMES)
def test_fk_against_golden(robot):
    g, d = load_golden(f"fk_{robot}"), get_desc(robot)
    sel = [d.link_names.index(n) for n in g["out_link_names"]]
    tree = co.Tree(d)
    assert rel_err(co.fk_forward(tree, g["q"], sel), g["H"]) < 1e-5
    assert rel_err(co.fk_backward(tree, g["q"], g["G"], sel), g["grad_q"]) < 1e-4


@pytest.mark.parametrize("case", ["panda_spheres3d", "panda_two_objects", "panda_boxes"])
def test_objective_against_golden_and_numpy(case):
    g, d = load_golden(f"planning_{case}"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    tree, cc = co.Tree(d), co.Cost(cost)
    q = g["q"].reshape(-1, 7)
    r = co.objective(tree, cc, q)
    for k, name in enumerate(["self", "obj", "ws", "ee"]):
        assert rel_err(r["terms"][:, k], g[f"cost_{name}"].reshape(-1)) < 1e-5
    cost.weights = np.array([1, 1, 1, 0], np.float32)
    r3 = co.objective(tree, co.Cost(cost), q)
    assert rel_err(r3["total"], g["cost_collision"].reshape(-1)) < 1e-5
    assert rel_err(r3["grad_q"], g["grad_collision"].reshape(-1, 7)) < 1e-4
    assert np.array_equal(r3["occ"] != 0, g["occ_all_default"].reshape(-1))
    r0 = co.objective(tree, co.Cost(cost), q, margin_override=0.0)
    assert np.array_equal(r0["occ"] != 0, g["occ_all_zero"].reshape(-1))


def test_c_and_numpy_agree_at_scale():
    """Same positions are not available to both (sin/cos differ by an ulp), so compare through the
    full path: values within tolerance, decisions equal except rare boundary cases."""
    g, d = load_golden("planning_panda_spheres3d"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    rng = np.random.default_rng(3)
    lo, hi = d.lo[d.controlled], d.hi[d.controlled]
    q = (lo + rng.random((50_000, 7)) * (hi - lo)).astype(np.float32)
    a = co.objective(co.Tree(d), co.Cost(cost), q)
    b = orc.objective(d, cost, q, dtype=np.float32)
    assert rel_err(a["terms"], b["terms"]) < 1e-5
    assert rel_err(a["grad_q"], b["grad_q"]) < 1e-4
    assert np.mean(a["occ"] != b["occ"]) < 1e-4
    assert np.mean(a["argmin"] != b["argmin"]) < 1e-4


def test_stretch_and_threads():
    g, d = load_golden("objective_stretch"), get_desc("stretch")
    r = co.objective(co.Tree(d), co.Cost(stretch_cost_from_golden(g)), g["q"])
    assert rel_err(r["total"], g["cost_total"]) < 1e-5
    assert rel_err(r["grad_q"], g["grad_total"]) < 1e-4
    assert co.max_threads() >= 1
