#!/usr/bin/env python
"""Panda with the 45-sphere per-link collision model (attached points -> runtime-tree kernels) and with the grasped
box: fused objective + gradient on EnvSpheres3D, 1,048,576 configurations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D  # noqa: E402
from torch_robotics_b200.environments.objects import GraspedObjectPandaBox  # noqa: E402
from torch_robotics_b200.robots.robot_panda import RobotPanda  # noqa: E402
from torch_robotics_b200.tasks.tasks import PlanningTask  # noqa: E402

TA = dict(device="cuda:0", dtype=torch.float32)
from torch_robotics_b200.environments.env_table_shelf import EnvTableShelf  # noqa: E402
envs = [("EnvSpheres3D", EnvSpheres3D(tensor_args=TA)), ("EnvTableShelf", EnvTableShelf(tensor_args=TA)),
        ("EnvTableShelf voxel SDF 2 cm", EnvTableShelf(precompute_sdf_obj_fixed=True, sdf_cell_size=0.02, tensor_args=TA))]
for env_name, env in envs:
  for name, robot in (("45-sphere model", RobotPanda(sphere_config=True, tensor_args=TA)),
                      ("grasped box (5 + 14 points, 66 pairs)", RobotPanda(grasped_object=GraspedObjectPandaBox(tensor_args=TA), tensor_args=TA))):
      task = PlanningTask(env=env, robot=robot, tensor_args=TA)
      q = robot.q_min + torch.rand(16384, 64, 7, device="cuda:0") * (robot.q_max - robot.q_min)
      for _ in range(3):
          task.compute_cost_and_grad(q)
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(5):
          task.compute_cost_and_grad(q)
      b.record()
      torch.cuda.synchronize()
      ms = a.elapsed_time(b) / 5
      print(json.dumps({"env": env_name, "model": name, "family": robot.diff_panda.device_model("cuda:0").kernel_family, "configs": q.shape[0] * q.shape[1],
                        "ms": ms, "configs_per_s": q.shape[0] * q.shape[1] / ms * 1e3}))
