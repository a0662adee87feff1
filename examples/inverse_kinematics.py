"""Batched Adam IK for the Panda - the B200 counterpart of the reference's examples/inverse_kinematics.py
(same target, same hyper-parameters), without the matplotlib part."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_robotics_b200.torch_kinematics_tree.geometrics.frame import Frame  # noqa: E402
from torch_robotics_b200.torch_kinematics_tree.geometrics.spatial_vector import y_rot, z_rot  # noqa: E402
from torch_robotics_b200.torch_kinematics_tree.geometrics.utils import SE3_distance  # noqa: E402
from torch_robotics_b200.torch_kinematics_tree.models.robots import DifferentiableFrankaPanda  # noqa: E402
from torch_robotics_b200.torch_utils.seed import fix_random_seed  # noqa: E402
from torch_robotics_b200.torch_utils.torch_timer import TimerCUDA  # noqa: E402

if __name__ == "__main__":
    fix_random_seed(0)
    batch_size = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    device = "cuda"
    tensor_args = dict(device=device, dtype=torch.float32)
    pos_target = torch.tensor([0.2, 0.4, 0.1], **tensor_args)
    rot_target = (z_rot(-torch.tensor(torch.pi / 2)) @ y_rot(-torch.tensor(torch.pi))).to(**tensor_args)
    H_target = Frame(rot=rot_target, trans=pos_target, device=device).get_transform_matrix()
    diff_panda = DifferentiableFrankaPanda(gripper=False, device=device)
    with TimerCUDA() as t:
        q_ik, idx_valid = diff_panda.inverse_kinematics(
            H_target, link_name="ee_link", batch_size=batch_size, max_iters=500, lr=2e-1, se3_eps=5e-2,
            eps_joint_lim=torch.pi / 64, print_freq=100, debug=False)
    print(f"\nIK time: {t.elapsed:.3f} sec")
    print(f"idx_valid: {idx_valid.nelement()}/{batch_size}")
    if idx_valid.nelement():
        H = diff_panda.compute_forward_kinematics_all_links(q_ik.detach()[idx_valid], link_list=["ee_link"]).squeeze(1)
        print("max SE3 error of the valid solutions:", float(SE3_distance(H, H_target).max()))
