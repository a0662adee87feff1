#!/usr/bin/env python
"""Per-call latency of the drop-in Python API at BASELINE config 3 (4096 x 64 Panda configurations): what a planner
pays per iteration for cost + gradient, next to the bare kernel time."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D  # noqa: E402
from torch_robotics_b200.robots.robot_panda import RobotPanda  # noqa: E402
from torch_robotics_b200.tasks.tasks import PlanningTask  # noqa: E402

TA = dict(device="cuda:0", dtype=torch.float32)
robot = RobotPanda(tensor_args=TA)
task = PlanningTask(env=EnvSpheres3D(tensor_args=TA), robot=robot, tensor_args=TA)
b, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 64)
q = (robot.q_min + torch.rand(b, h, 7, device="cuda:0") * (robot.q_max - robot.q_min)).requires_grad_(True)


def step():
    q.grad = None
    c = task.compute_collision_cost(q)
    c.sum().backward()
    return c


def wall(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


us_api = wall(step)
us_fwd = wall(lambda: task.compute_collision_cost(q.detach()))
us_mask = wall(lambda: task.compute_collision(q.detach()))
dc = task._fused_cost(q.device)
q2 = q.detach().reshape(-1, 7)
tot, grad = torch.empty(b * h, device="cuda:0"), torch.empty(b * h, 7, device="cuda:0")
us_direct = wall(lambda: task.compute_cost_and_grad(q))
us_raw = wall(lambda: dc.objective(q2, want_terms=False, out_total=tot, out_grad=grad))
print(json.dumps({"configs": b * h, "us_cost_and_backward_api": us_api, "us_cost_only_api": us_fwd, "us_collision_mask_api": us_mask,
                  "us_cost_and_grad_direct_api": us_direct, "us_raw_objective_call": us_raw, "configs_per_s_api": b * h / us_api * 1e6}))
