"""TEST / BENCH INFRASTRUCTURE - never imported by the product package.

Runs the UNMODIFIED reference (anindex/torch_robotics) through its own public API for BASELINE config 3
(Panda planning objective: `PlanningTask.compute_collision_cost` = self + object + workspace terms, tasks.py:135, plus
`EESE3DistanceField`, distance_fields.py:335-356, forward and `cost.sum().backward()`).  Used only by
`bench.py --impl reference` and by the `cpu_baseline` leg of `bench.py`.

The reference is looked up at /root/reference (build container) and then at baseline/_ref (the offline
`pip install --target baseline/_ref` copy that travels to the GPU box).  It imports `urdf_parser_py` and `matplotlib`,
which are not installed in this image: tests/golden/_shims holds the stand-ins (the same ones the golden vectors were
generated with).  Nothing here touches the CUDA library.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "tests", "golden", "_shims")


def locate_reference():
    """Path that holds the reference's `torch_robotics` package, or None."""
    for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(p, "torch_robotics", "__init__.py")) or os.path.isdir(os.path.join(p, "torch_robotics", "tasks")):
            return p
    return None


class ReferenceObjective:
    """BASELINE config 3 through the reference's classes on `device` ('cpu' or 'cuda:0')."""

    def __init__(self, device="cpu", threads=None):
        import torch
        self.torch = torch
        self.path = locate_reference()
        if self.path is None:
            raise ImportError("the reference is neither at /root/reference nor at baseline/_ref")
        if threads:
            torch.set_num_threads(int(threads))
        for p in (self.path, SHIMS):
            if p not in sys.path:
                sys.path.insert(0, p)
        with contextlib.redirect_stdout(io.StringIO()):
            from torch_robotics.environments.env_spheres_3d import EnvSpheres3D
            from torch_robotics.robots.robot_panda import RobotPanda
            from torch_robotics.tasks.tasks import PlanningTask
            from torch_robotics.torch_kinematics_tree.geometrics.frame import Frame
            from torch_robotics.torch_kinematics_tree.geometrics.spatial_vector import y_rot, z_rot
            from torch_robotics.torch_planning_objectives.fields.distance_fields import EESE3DistanceField
            ta = dict(device=device, dtype=torch.float32)
            self.robot = RobotPanda(tensor_args=ta)
            self.task = PlanningTask(env=EnvSpheres3D(tensor_args=ta), robot=self.robot, tensor_args=ta)
            # target of examples/inverse_kinematics.py:27-30
            rot = (z_rot(-torch.tensor(torch.pi / 2)) @ y_rot(-torch.tensor(torch.pi))).to(**ta)
            H = Frame(rot=rot, trans=torch.tensor([0.2, 0.4, 0.1], **ta), device=device).get_transform_matrix()
            self.ee = EESE3DistanceField(H, w_pos=1.0, w_rot=1.0, square=True)
        self.device = device

    def step(self, q):
        """q [b, h, 7] on self.device -> (cost [b, h], grad_q [b, h, 7]); one forward + backward of the objective."""
        b, h, d = q.shape
        qq = q.detach().clone().requires_grad_(True)
        cost = self.task.compute_collision_cost(qq)
        Hs = self.robot.diff_panda.compute_forward_kinematics_all_links(qq.reshape(-1, d))
        cost = cost + self.ee.compute_costs_impl(qq, Hs).reshape(b, h)
        cost.sum().backward()
        return cost.detach(), qq.grad

    def limits(self):
        return self.robot.q_min.cpu().numpy(), self.robot.q_max.cpu().numpy()


def time_steps(obj: ReferenceObjective, q, steps: int, warmup: int):
    """Wall-clock seconds per step (mean over `steps` after `warmup`), synchronising CUDA when the reference runs there."""
    torch = obj.torch
    sync = (lambda: torch.cuda.synchronize()) if str(obj.device).startswith("cuda") else (lambda: None)
    for _ in range(warmup):
        obj.step(q)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        obj.step(q)
    sync()
    return (time.perf_counter() - t0) / max(steps, 1)
