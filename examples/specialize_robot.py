"""A robot outside the shipped set on the unrolled kernels: compile it once (about half a minute of nvcc, cached by
content), then FK / its reverse mode / the fused planning objective of this tree run on kernels generated for it.

    python examples/specialize_robot.py            (needs a B200 and nvcc)
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_robotics_b200.torch_kinematics_tree.models.robot_tree import DifferentiableTree  # noqa: E402
from torch_robotics_b200.torch_kinematics_tree.utils.files import get_robot_path  # noqa: E402

device = "cuda:0"
urdf = os.path.join(str(get_robot_path()), "kuka_iiwa", "urdf", "iiwa7_allegro.urdf")      # iiwa7 arm + Allegro hand: 29 links, 23 DoF
tree = DifferentiableTree(urdf, "iiwa7_allegro", device=device)
q = torch.rand(1 << 18, tree._n_dofs, device=device) - 0.5


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("kernel family before:", tree.device_model(device).kernel_family)
ms0 = timed(lambda: tree.compute_forward_kinematics_all_links(q))
H0 = tree.compute_forward_kinematics_all_links(q[:1024])

t0 = time.perf_counter()
path = tree.specialize()
print(f"specialised library: {os.path.basename(path)} ({time.perf_counter() - t0:.1f} s; cached for the next process)")
print("kernel family after: ", tree.device_model(device).kernel_family)
ms1 = timed(lambda: tree.compute_forward_kinematics_all_links(q))
H1 = tree.compute_forward_kinematics_all_links(q[:1024])
print(f"FK of {q.shape[0]} configurations: {ms0:.3f} ms -> {ms1:.3f} ms; max |dH| = {(H1 - H0).abs().max().item():.1e}")
