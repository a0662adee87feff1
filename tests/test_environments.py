"""The reference's Panda 3-D environments as shipped here (environments/env_table_shelf.py, env_maze_boxes_3d.py,
env_spheres_3d_extra_objects.py): obstacle tables bit-identical to the reference's, and the planning objective on them
equal to the reference's (tests/golden/planning_panda_{table_shelf,maze_boxes,spheres3d_extra}.npz)."""
import numpy as np
import pytest
import torch

from conftest import get_desc, load_golden, planning_cost_from_golden, rel_err
from oracle import c_oracle as co
from oracle import restatement as orc

ENVS = ["table_shelf", "maze_boxes", "spheres3d_extra"]


def make_env(tag, tensor_args):
    from torch_robotics_b200.environments.env_maze_boxes_3d import EnvMazeBoxes3D
    from torch_robotics_b200.environments.env_spheres_3d_extra_objects import EnvSpheres3DExtraObjects
    from torch_robotics_b200.environments.env_table_shelf import EnvTableShelf
    return dict(table_shelf=EnvTableShelf, maze_boxes=EnvMazeBoxes3D, spheres3d_extra=EnvSpheres3DExtraObjects)[tag](tensor_args=tensor_args)


@pytest.mark.parametrize("tag", ENVS)
def test_environment_tables_bit_identical_to_reference(tag):
    g = load_golden(f"planning_panda_{tag}")
    env = make_env(tag, dict(device="cpu", dtype=torch.float32))
    assert np.array_equal(env.limits.numpy(), g["env_limits"])
    objs = env.get_df_obj_list()
    assert len(objs) == int(g["n_obj"])
    for oi, obj in enumerate(objs):
        centers, radii, pos, ori, boxes = obj.as_sphere_table()
        assert np.array_equal(centers, g[f"obj{oi}_centers"].reshape(-1, 3)) and np.array_equal(radii, g[f"obj{oi}_radii"])
        assert np.array_equal(pos, g[f"obj{oi}_pos"]) and np.array_equal(ori, g[f"obj{oi}_ori_wxyz"])
        assert len(boxes) == int(g[f"obj{oi}_n_box_fields"])
        for bi, (bc, bs, rounded) in enumerate(boxes):
            assert np.array_equal(bc, g[f"obj{oi}_box{bi}_centers"]) and np.array_equal(bs, g[f"obj{oi}_box{bi}_sizes"])
            assert bool(rounded) == bool(g[f"obj{oi}_box{bi}_rounded"])


@pytest.mark.parametrize("tag", ENVS)
def test_oracles_match_reference(tag):
    g, d = load_golden(f"planning_panda_{tag}"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    q = g["q"].reshape(-1, 7)
    for o in (orc.objective(d, cost, q, dtype=np.float32), co.objective(co.Tree(d), co.Cost(cost), q)):
        for k, name in enumerate(("self", "obj", "ws", "ee")):
            assert rel_err(o["terms"][:, k], g[f"cost_{name}"].reshape(-1)) < 1e-5
        assert np.mean((o["occ"] != 0) != g["occ_all_default"].reshape(-1)) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ENVS)
def test_gpu_planning_task_matches_reference(tag):
    from torch_robotics_b200.robots.robot_panda import RobotPanda
    from torch_robotics_b200.tasks.tasks import PlanningTask
    TA = dict(device="cuda:0", dtype=torch.float32)
    g = load_golden(f"planning_panda_{tag}")
    task = PlanningTask(env=make_env(tag, TA), robot=RobotPanda(tensor_args=TA), tensor_args=TA)
    q = torch.tensor(g["q"], device="cuda:0").requires_grad_(True)
    c = task.compute_collision_cost(q)
    assert rel_err(c.detach().cpu().numpy(), g["cost_collision"]) < 1e-5
    c.sum().backward()
    assert rel_err(q.grad.cpu().numpy(), g["grad_collision"]) < 1e-4
    with torch.no_grad():
        pos_ref = torch.tensor(g["link_pos"], device="cuda:0")      # stage-wise: masks on the reference's own positions
        for mtag, kw in (("default", {}), ("zero", {"margin": 0.0})):
            occ = task.df_collision_objects.compute_cost(q, pos_ref, field_type="occupancy", **kw)
            assert np.array_equal(occ.cpu().numpy(), g[f"occ_obj_{mtag}"])
        assert np.mean(task.compute_collision(q).cpu().numpy() != g["occ_all_default"]) < 0.02
    # the same environment through its precomputed voxel SDF: costs agree to the grid resolution (five points, each
    # off by up to ~2.5 cell diagonals: the reference samples cell i at lo + i*extent/(n-1) but bins with extent/n)
    env_grid = type(task.env)(precompute_sdf_obj_fixed=True, sdf_cell_size=0.02, tensor_args=TA)
    task_grid = PlanningTask(env=env_grid, robot=task.robot, tensor_args=TA)
    cg = task_grid.compute_collision_cost(q.detach())
    assert torch.isfinite(cg).all() and (cg - c.detach()).abs().max().item() < 5 * 2.5 * 0.02 * 3 ** 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["table_shelf", "maze_boxes", "spheres3d_extra", "boxes", "two_objects"])
@pytest.mark.parametrize("with_ee", [True, False])
def test_gpu_general_scene_headline_kernel_vs_oracle(tag, with_ee):
    """Scenes beyond the single sphere object - posed objects of rounded boxes, several sphere objects, spheres and boxes mixed,
    unequal radii - run on k_objective_hot<..., SCENE = 1>, with or without an EE term: terms, gradients, occupancy (default
    margins and margin = 0) and the per-object arg-min against the C oracle on random configurations; the fused block sums exist
    (only the headline kernel has them)."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    g, d = load_golden(f"planning_panda_{tag}"), get_desc("panda")
    cost = planning_cost_from_golden(g, with_ee=with_ee)
    m = DeviceModel(d, "cuda:0")
    dc = DeviceCost(m, cost)
    rng = np.random.default_rng(3)
    lo, hi = d.lo[d.controlled], d.hi[d.controlled]
    q = (lo + rng.random((40_009, 7), dtype=np.float32) * (hi - lo)).astype(np.float32)
    gq = g["q"].reshape(-1, 7)[:64]
    q[:len(gq)] = gq
    qt = torch.as_tensor(q, device="cuda:0")
    r = dc.objective(qt, want_terms=True, want_occupancy=True, want_argmin=True)
    o = co.objective(co.Tree(d), co.Cost(cost), q)
    assert rel_err(r["terms"].cpu().numpy(), o["terms"]) < 1e-5
    same = np.all((r["argmin"].cpu().numpy() == o["argmin"]).reshape(len(q), -1), axis=1)
    assert np.mean(same) > 0.999
    gd = np.abs(r["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    assert np.all(gd[same] < 1e-4) and np.mean(gd < 1e-4) > 0.999
    assert np.mean(r["occupancy"].cpu().numpy() == o["occ"]) > 0.9995
    r0 = dc.objective(qt, want_grad=False, want_terms=False, want_occupancy=True, margin_override=0.0)
    o0 = co.objective(co.Tree(d), co.Cost(cost), q, margin_override=0.0)
    assert np.mean(r0["occupancy"].cpu().numpy() == o0["occ"]) > 0.9995
    assert torch.equal(r0["total"], r["total"])
    total, grad = torch.empty(len(q), device="cuda:0"), torch.empty(len(q), 7, device="cuda:0")
    parts = dc.objective_sum(qt, out_grad=grad, out_total=total)
    assert torch.equal(total, r["total"]) and torch.equal(grad, r["grad_q"])
    assert abs(float(parts.sum(dtype=torch.float64)) - float(total.sum(dtype=torch.float64))) < 1e-3 * abs(float(total.sum(dtype=torch.float64)))
