#!/usr/bin/env python
"""Fused Panda objective (cost + gradient, PlanningTask.compute_cost_and_grad) on the shipped 3-D environments,
1,048,576 configurations each: which objective-kernel variant serves them and how fast."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_robotics_b200.environments.env_maze_boxes_3d import EnvMazeBoxes3D  # noqa: E402
from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D  # noqa: E402
from torch_robotics_b200.environments.env_spheres_3d_extra_objects import EnvSpheres3DExtraObjects  # noqa: E402
from torch_robotics_b200.environments.env_table_shelf import EnvTableShelf  # noqa: E402
from torch_robotics_b200.robots.robot_panda import RobotPanda  # noqa: E402
from torch_robotics_b200.tasks.tasks import PlanningTask  # noqa: E402

TA = dict(device="cuda:0", dtype=torch.float32)
robot = RobotPanda(tensor_args=TA)
b, h = 16384, 64
q = (robot.q_min + torch.rand(b, h, 7, device="cuda:0") * (robot.q_max - robot.q_min))
cases = [("EnvSpheres3D", lambda: EnvSpheres3D(tensor_args=TA)),
         ("EnvSpheres3DExtraObjects", lambda: EnvSpheres3DExtraObjects(tensor_args=TA)),
         ("EnvTableShelf", lambda: EnvTableShelf(tensor_args=TA)),
         ("EnvMazeBoxes3D", lambda: EnvMazeBoxes3D(tensor_args=TA)),
         ("EnvSpheres3D voxel SDF 1 cm", lambda: EnvSpheres3D(precompute_sdf_obj_fixed=True, sdf_cell_size=0.01, tensor_args=TA)),
         ("EnvTableShelf voxel SDF 1 cm", lambda: EnvTableShelf(precompute_sdf_obj_fixed=True, sdf_cell_size=0.01, tensor_args=TA))]
for name, mk in cases:
    task = PlanningTask(env=mk(), robot=robot, tensor_args=TA)
    for _ in range(3):
        task.compute_cost_and_grad(q)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        task.compute_cost_and_grad(q)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 10
    print(json.dumps({"env": name, "configs": b * h, "ms": ms, "configs_per_s": b * h / ms * 1e3}))
