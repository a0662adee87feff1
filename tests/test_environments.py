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
