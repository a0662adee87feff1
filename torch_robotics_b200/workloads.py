"""Benchmark / smoke workloads assembled through the public drop-in classes."""
import torch

from .environments.env_spheres_3d import EnvSpheres3D
from .robots.robot_panda import RobotPanda
from .tasks.tasks import PlanningTask
from .torch_kinematics_tree.geometrics.spatial_vector import y_rot, z_rot
from .torch_planning_objectives.fields.distance_fields import EESE3DistanceField


def ik_example_target():
    """EE target of the reference's IK example (examples/inverse_kinematics.py:27-30)."""
    H = torch.eye(4)
    H[:3, :3] = (z_rot(-torch.tensor(torch.pi / 2)) @ y_rot(-torch.tensor(torch.pi)))[0]
    H[:3, 3] = torch.tensor([0.2, 0.4, 0.1])
    return H


def panda_spheres3d_objective(device="cpu"):
    """(TreeDescriptor, CostDescriptor) of BASELINE config 3: RobotPanda + EnvSpheres3D + default
    PlanningTask + EESE3DistanceField(w_pos=1, w_rot=1, square=True) on ee_link."""
    ta = dict(device=device, dtype=torch.float32)
    robot = RobotPanda(tensor_args=ta)
    task = PlanningTask(env=EnvSpheres3D(tensor_args=ta), robot=robot, tensor_args=ta)
    ee = EESE3DistanceField(ik_example_target(), w_pos=1.0, w_rot=1.0, square=True)
    return robot.diff_panda._desc, task.cost_descriptor(ee)
