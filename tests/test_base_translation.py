"""A base pose that is a pure translation (DifferentiableTree.update_base_pose, reference robot_tree.py:133): PlanningTask
evaluates the fused objective in the robot's base frame on the scene shifted by -t, so that the tree keeps the unrolled
kernels (descriptor.cost_in_translated_frame, tasks/tasks.py:_fused_cost).  CPU: the shifted descriptor gives the world-frame
objective (C oracle with the base pose on the tree vs the oracle on the shifted scene).  GPU: PlanningTask on a translated Panda
against the oracle with the base pose, on the compiled-in family."""
import copy

import numpy as np
import pytest

from conftest import get_desc, load_golden, planning_cost_from_golden, rel_err
from oracle import c_oracle as co

T_BASE = np.array([0.31, -0.17, 0.12], np.float32)


def _random_q(d, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = d.lo[d.controlled], d.hi[d.controlled]
    return (lo + rng.random((n, d.n_dofs), dtype=np.float32) * (hi - lo)).astype(np.float32)


@pytest.mark.parametrize("case", ["spheres3d", "table_shelf", "boxes", "two_objects"])
def test_shifted_scene_gives_the_world_frame_objective(case):
    from torch_robotics_b200.descriptor import cost_in_translated_frame
    d = get_desc("panda")
    cost = planning_cost_from_golden(load_golden(f"planning_panda_{case}"))
    db = copy.copy(d)
    db.base_t = T_BASE.copy()
    q = _random_q(d, 20_000, 1)
    world = co.objective(co.Tree(db), co.Cost(cost), q)
    shifted = cost_in_translated_frame(cost, T_BASE)
    base = co.objective(co.Tree(d), co.Cost(shifted), q)
    assert not np.allclose(world["terms"], co.objective(co.Tree(d), co.Cost(cost), q)["terms"])      # the base pose matters
    assert rel_err(base["terms"], world["terms"]) < 1e-5
    same = np.all((base["argmin"] == world["argmin"]).reshape(len(q), -1), axis=1)
    assert np.mean(same) > 0.999
    gd = np.abs(base["grad_q"] - world["grad_q"]).max(axis=1) / np.abs(world["grad_q"]).max()
    assert np.all(gd[same] < 1e-4)
    assert np.mean(base["occ"] == world["occ"]) > 0.9995
    # the input descriptor is untouched, identity-pose objects stay identity-pose objects
    assert np.array_equal(cost.obj_identity, shifted.obj_identity) and not np.shares_memory(cost.sph_center, shifted.sph_center)


def test_base_translation_detection():
    from torch_robotics_b200.torch_kinematics_tree.models.robots import DifferentiableFrankaPanda
    tree = DifferentiableFrankaPanda(gripper=False, device="cuda:0")
    assert tree.base_translation() is None
    tree.update_base_pose([0.3, 0.0, -0.1, 1.0, 0.0, 0.0, 0.0])
    assert np.array_equal(tree.base_translation(), np.array([0.3, 0.0, -0.1], np.float32))
    tree.update_base_pose([0.3, 0.0, -0.1, 0.9238795, 0.0, 0.0, 0.3826834])      # rotated about z: no pull-back
    assert tree.base_translation() is None


@pytest.mark.gpu
def test_gpu_planning_task_with_a_translated_base_runs_the_compiled_in_kernels():
    import torch
    from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D
    from torch_robotics_b200.robots.robot_panda import RobotPanda
    from torch_robotics_b200.tasks.tasks import PlanningTask
    TA = dict(device="cuda:0", dtype=torch.float32)
    robot = RobotPanda(tensor_args=TA)
    task = PlanningTask(env=EnvSpheres3D(tensor_args=TA), robot=robot, tensor_args=TA)
    d = robot.diff_panda._desc
    q = _random_q(d, 8192, 2)
    qt = torch.as_tensor(q, device="cuda:0").reshape(128, 64, 7)
    c0, _ = task.compute_cost_and_grad(qt)
    robot.diff_panda.update_base_pose([float(T_BASE[0]), float(T_BASE[1]), float(T_BASE[2]), 1.0, 0.0, 0.0, 0.0])
    c1, g1 = task.compute_cost_and_grad(qt)
    occ1 = task.compute_collision(qt)
    assert robot.diff_panda.device_model("cuda:0").kernel_family == "generic"                     # FK of the posed tree
    assert robot.diff_panda.device_model_in_base_frame("cuda:0").kernel_family == "panda"          # the objective's tree
    assert not torch.allclose(c0, c1)
    world = co.objective(co.Tree(robot.diff_panda._desc), co.Cost(task.cost_descriptor()), q)
    assert rel_err(c1.reshape(-1).cpu().numpy(), world["total"]) < 1e-5
    gd = np.abs(g1.reshape(-1, 7).cpu().numpy() - world["grad_q"]).max(axis=1) / np.abs(world["grad_q"]).max()
    assert np.mean(gd < 1e-4) > 0.999
    assert np.mean(occ1.reshape(-1).cpu().numpy() == (world["occ"] != 0)) > 0.9995
