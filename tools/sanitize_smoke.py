#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import get_desc, load_golden, planning_cost_from_golden, stretch_cost_from_golden
from torch_robotics_b200._lib import DeviceCost, DeviceModel
from torch_robotics_b200.descriptor import compile_urdf

dev = "cuda:0"
rng = np.random.default_rng(0)
for robot, n in (("panda", 1000), ("stretch", 333), ("shadow_hand", 97)):
    d = get_desc(robot)
    m = DeviceModel(d, dev)
    q = torch.as_tensor(rng.uniform(-3, 3, (n, d.n_dofs)).astype(np.float32), device=dev)
    H = m.fk_forward(q)                                   # TMA bulk-store kernel
    sel = [int(v) for v in rng.permutation(d.n_links)[:3]]
    Hs = m.fk_forward(q, sel)                             # staged-store kernel
    g = m.fk_backward(q, torch.randn_like(H))             # cp.async double-buffered kernel
    gs = m.fk_backward(q, torch.randn_like(Hs), sel)
# a URDF outside the compiled-in set -> runtime-tree family
import tempfile
urdf = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torch_robotics_b200", "data", "urdf", "robots", "kuka_iiwa", "urdf", "iiwa7_allegro.urdf")
dg = compile_urdf(urdf, "iiwa7_allegro")
mg = DeviceModel(dg, dev)
assert mg.kernel_family == "generic"
qg = torch.as_tensor(rng.uniform(-1, 1, (515, dg.n_dofs)).astype(np.float32), device=dev)
Hg = mg.fk_forward(qg); gg = mg.fk_backward(qg, torch.randn_like(Hg))
for case, robot in (("planning_panda_spheres3d", "panda"), ("planning_panda_two_objects", "panda"), ("planning_panda_boxes", "panda")):
    g = load_golden(case); d = get_desc(robot); m = DeviceModel(d, dev)
    dc = DeviceCost(m, planning_cost_from_golden(g))
    q = torch.as_tensor(g["q"].reshape(-1, 7), device=dev)
    r = dc.objective(q, want_terms=True, want_occupancy=True, want_argmin=True)
    H = m.fk_forward(q)
    r2 = dc.cost_stage(H[:, :, :3, 3].contiguous(), H[:, 10].contiguous(), want_occupancy=True, want_argmin=True)
g = load_golden("objective_stretch"); d = get_desc("stretch"); m = DeviceModel(d, dev)
DeviceCost(m, stretch_cost_from_golden(g)).objective(torch.as_tensor(g["q"], device=dev))
g = load_golden("ik_panda"); m = DeviceModel(get_desc("panda"), dev)
m.ik_adam(torch.as_tensor(g["q0"], device=dev), g["H_target"], 10, g["lower"], g["upper"], n_iters=5, lr=0.2, se3_eps=0.05, want_loss_stats=True)
# runtime-tree family on the compiled-in robots too (objective with and without gradient, sub-selections)
os.environ["TRK_B200_FORCE_GENERIC"] = "1"
for robot, n in (("panda", 517), ("stretch", 131), ("shadow_hand", 65)):
    d = get_desc(robot); mg2 = DeviceModel(d, dev)
    assert mg2.kernel_family == "generic"
    q = torch.as_tensor(rng.uniform(-3, 3, (n, d.n_dofs)).astype(np.float32), device=dev)
    H = mg2.fk_forward(q); mg2.fk_backward(q, torch.randn_like(H))
    sel = [int(v) for v in rng.permutation(d.n_links)[:3]]
    Hs = mg2.fk_forward(q, sel); mg2.fk_backward(q, torch.randn_like(Hs), sel)
    mg2.link_jacobian(q)
gq = load_golden("planning_panda_boxes"); mgp = DeviceModel(get_desc("panda"), dev)
dcg = DeviceCost(mgp, planning_cost_from_golden(gq)); qq = torch.as_tensor(gq["q"].reshape(-1, 7), device=dev)
dcg.objective(qq, want_terms=True, want_occupancy=True, want_argmin=True); dcg.objective(qq, want_grad=False)
DeviceCost(DeviceModel(get_desc("stretch"), dev), stretch_cost_from_golden(load_golden("objective_stretch"))).objective(torch.as_tensor(load_golden("objective_stretch")["q"], device=dev))
del os.environ["TRK_B200_FORCE_GENERIC"]
# attached collision points (45-sphere Panda model, grasped box) and the voxel SDF (precompute + lookup + standalone query)
from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D
from torch_robotics_b200.environments.objects import GraspedObjectPandaBox
from torch_robotics_b200.robots.robot_panda import RobotPanda
from torch_robotics_b200.tasks.tasks import PlanningTask
TA = dict(device=dev, dtype=torch.float32)
env = EnvSpheres3D(precompute_sdf_obj_fixed=True, sdf_cell_size=0.1, tensor_args=TA)
for robot in (RobotPanda(tensor_args=TA), RobotPanda(sphere_config=True, tensor_args=TA),
              RobotPanda(grasped_object=GraspedObjectPandaBox(tensor_args=TA), tensor_args=TA)):
    task = PlanningTask(env=env, robot=robot, tensor_args=TA)
    qt = (robot.q_min + torch.rand(9, 7, 7, device=dev) * (robot.q_max - robot.q_min)).requires_grad_(True)
    task.compute_collision_cost(qt).sum().backward()
    task.compute_collision(qt.detach())
    robot.diff_panda.compute_analytical_jacobian_all_links(qt.detach().reshape(-1, 7))
# round 2 kernels: attached-body objective (plain sphere scene), fused cost sums, IK with the per-block top-k, geometric Jacobian
env2 = EnvSpheres3D(tensor_args=TA)
for robot in (RobotPanda(sphere_config=True, tensor_args=TA), RobotPanda(grasped_object=GraspedObjectPandaBox(tensor_args=TA), tensor_args=TA)):
    task = PlanningTask(env=env2, robot=robot, tensor_args=TA)
    qt = (robot.q_min + torch.rand(5, 31, 7, device=dev) * (robot.q_max - robot.q_min))
    task.compute_cost_and_grad(qt)
    task.compute_collision(qt, margin=0.0)
g3 = load_golden("planning_panda_spheres3d"); m3 = DeviceModel(get_desc("panda"), dev)
dc3 = DeviceCost(m3, planning_cost_from_golden(g3))
q3 = torch.as_tensor(rng.uniform(-2, 2, (777, 7)).astype(np.float32), device=dev)
dc3.objective_sum(q3, out_grad=torch.empty_like(q3), out_total=torch.empty(777, device=dev))
dc3.objective(q3, want_grad=False, want_terms=True, want_occupancy=True, want_argmin=True)
gi = load_golden("ik_panda")
m3.ik_adam(torch.as_tensor(rng.uniform(-1, 1, (333, 7)).astype(np.float32), device=dev), gi["H_target"], 10, gi["lower"], gi["upper"], n_iters=7, lr=0.2,
           se3_eps=0.05, best_k=8)
m3.geometric_jacobian(q3, 10, 0x7f)
DeviceModel(get_desc("tiago_dual_holo"), dev).geometric_jacobian(torch.zeros(40, 14, device=dev), 12, 0x3fff)
# later in round 2: the headline kernel's scene variants (no EE term; boxes, two sphere objects, voxel grid + extra object),
# and a library specialised at run time
from torch_robotics_b200.environments.env_maze_boxes_3d import EnvMazeBoxes3D
from torch_robotics_b200.environments.env_spheres_3d_extra_objects import EnvSpheres3DExtraObjects
from torch_robotics_b200.environments.env_table_shelf import EnvTableShelf
robot_p = RobotPanda(tensor_args=TA)
for mk in (lambda: EnvSpheres3D(tensor_args=TA), lambda: EnvTableShelf(tensor_args=TA), lambda: EnvMazeBoxes3D(tensor_args=TA),
           lambda: EnvSpheres3DExtraObjects(tensor_args=TA),
           lambda: EnvSpheres3DExtraObjects(precompute_sdf_obj_fixed=True, sdf_cell_size=0.05, tensor_args=TA)):
    task = PlanningTask(env=mk(), robot=robot_p, tensor_args=TA)
    qt = (robot_p.q_min + torch.rand(3, 43, 7, device=dev) * (robot_p.q_max - robot_p.q_min))
    task.compute_cost_and_grad(qt)
    task.compute_collision(qt, margin=0.0)
    dcs = task._fused_cost(qt.device)
    dcs.objective(qt.reshape(-1, 7), want_grad=True, want_terms=True, want_occupancy=True, want_argmin=True)

for robot_a in (RobotPanda(sphere_config=True, tensor_args=TA), RobotPanda(grasped_object=GraspedObjectPandaBox(tensor_args=TA), tensor_args=TA)):
    for mk in (lambda: EnvTableShelf(tensor_args=TA), lambda: EnvSpheres3DExtraObjects(precompute_sdf_obj_fixed=True, sdf_cell_size=0.05, tensor_args=TA)):
        task = PlanningTask(env=mk(), robot=robot_a, tensor_args=TA)
        qt = (robot_a.q_min + torch.rand(2, 37, 7, device=dev) * (robot_a.q_max - robot_a.q_min))
        task.compute_cost_and_grad(qt)
        task._fused_cost(qt.device).objective(qt.reshape(-1, 7), want_grad=False, want_terms=True, want_occupancy=True, want_argmin=True, margin_override=0.0)
import specialized_cases as sc
from torch_robotics_b200 import specialize as sp
d_i = sc.iiwa7_allegro_desc()
m_i = DeviceModel(d_i, dev, library=sp.build_specialized_library(d_i))
q_i = torch.rand(77, d_i.n_dofs, device=dev) - 0.5
H_i = m_i.fk_forward(q_i)
m_i.fk_backward(q_i, torch.randn_like(H_i))
m_i.link_jacobian(q_i)
X = (torch.rand(300, 3, device=dev) * 2.4 - 1.2).requires_grad_(True)
env.grid_map_sdf_obj_fixed(X).sum().backward()
torch.cuda.synchronize()
print("sanitize_smoke: all kernel families ran")
