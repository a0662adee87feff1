"""RobotPanda: Franka Panda (no gripper) with the reference's collision model
(robots/robot_panda.py:21-184): five object-collision spheres and ten self-collision link pairs."""
from collections import OrderedDict

import numpy as np
import torch

from ..torch_kinematics_tree.geometrics.utils import link_pos_from_link_tensor, link_rot_from_link_tensor
from ..torch_kinematics_tree.models.robots import DifferentiableFrankaPanda
from .robot_base import RobotBase

# link -> radius of its collision sphere (robot_panda.py:47-71)
OBJECT_COLLISION_SPHERES = OrderedDict([
    ("panda_link2", 0.125), ("panda_link3", 0.125), ("panda_link5", 0.13), ("panda_link7", 0.1), ("panda_hand", 0.08)])
# link -> links it is checked against (robot_panda.py:81-86)
SELF_COLLISION_PAIRS = OrderedDict([
    ("panda_link4", ["panda_link1"]),
    ("panda_link5", ["panda_link0", "panda_link1", "panda_link2"]),
    ("panda_link6", ["panda_link0", "panda_link1", "panda_link2"]),
    ("panda_hand", ["panda_link0", "panda_link1", "panda_link2"])])
GRASPED_OBJECT_SELF_LINKS = ["panda_link0", "panda_link1", "panda_link2", "panda_link3"]


class RobotPanda(RobotBase):
    """`sphere_config`: path of a *_sphere_config.json (or True for the shipped Panda file) switches the
    object-collision model from the reference's five link-origin spheres to the per-link multi-sphere model of
    data/configs/panda (45 spheres with local offsets), realised as points attached to the links."""

    def __init__(self, use_self_collision_storm=False, grasped_object=None, tensor_args=None, sphere_config=None, **kwargs):
        if use_self_collision_storm:
            raise NotImplementedError("the STORM learned self-collision model needs storm_kit weights (out of scope)")
        self.gripper = False
        self.link_name_ee = "ee_link"
        self.link_name_grasped_object = "grasped_object"
        self.diff_panda = DifferentiableFrankaPanda(gripper=False, device=tensor_args["device"], grasped_object=grasped_object)
        self.diff_model = self.diff_panda
        self.jl_lower, self.jl_upper, _, _ = self.diff_panda.get_joint_limit_array()
        q_limits = torch.tensor(np.array([self.jl_lower, self.jl_upper]), **tensor_args)
        name_to_idx = self.diff_panda._name_to_idx_map

        obj_links = list(OBJECT_COLLISION_SPHERES)
        obj_radii = list(OBJECT_COLLISION_SPHERES.values())
        obj_idxs = [name_to_idx[n] for n in obj_links]
        if sphere_config:
            from ..descriptor import load_sphere_config
            from ..torch_kinematics_tree.utils.files import get_configs_path
            path = get_configs_path() / "panda" / "panda_sphere_config.json" if sphere_config is True else sphere_config
            spheres = [s for s in load_sphere_config(path) if s[0] in name_to_idx]
            obj_links, obj_idxs = self.diff_panda.attach_points([(s[0], s[1]) for s in spheres])
            obj_radii = [s[2] for s in spheres]
        grasped_idxs = None
        if grasped_object is not None:      # the 14 box points ride on the `grasped_object` link (robot_panda.py:151-159)
            pts = grasped_object.base_points_for_collision.detach().cpu().tolist()
            _, grasped_idxs = self.diff_panda.attach_points([(self.link_name_grasped_object, p) for p in pts])
        self_links = set(GRASPED_OBJECT_SELF_LINKS)
        for k, v in SELF_COLLISION_PAIRS.items():
            self_links.add(k)
            self_links.update(v)
        self_links = sorted(self_links)     # name order, as the reference sorts them (robot_panda.py:101)
        super().__init__(
            name="RobotPanda", q_limits=q_limits, grasped_object=grasped_object,
            link_names_for_object_collision_checking=obj_links,
            link_margins_for_object_collision_checking=obj_radii,
            link_idxs_for_object_collision_checking=obj_idxs, grasped_object_point_idxs=grasped_idxs,
            margin_for_grasped_object_collision_checking=0.001,
            num_interpolated_points_for_object_collision_checking=len(obj_links),
            link_names_for_self_collision_checking=self_links,
            link_names_pairs_for_self_collision_checking=SELF_COLLISION_PAIRS,
            link_idxs_for_self_collision_checking=[name_to_idx[n] for n in self_links],
            num_interpolated_points_for_self_collision_checking=len(self_links),
            self_collision_margin_robot=0.05,
            link_names_for_self_collision_checking_with_grasped_object=GRASPED_OBJECT_SELF_LINKS,
            self_collision_margin_grasped_object=0.05, tensor_args=tensor_args, **kwargs)

    def fk_map_collision_impl(self, q, **kwargs):
        """q [b,h,7] | [b,7] -> positions of all links, then of the attached points (grasped object, multi-sphere
        model) [b,h,11(+P),3] | [b,11(+P),3] (robot_panda.py:138-170)."""
        shp = q.shape
        if q.ndim == 3:
            q = q.reshape(-1, shp[-1])
        elif q.ndim != 2:
            raise NotImplementedError
        H = self.diff_panda.compute_forward_kinematics_all_links(q, link_list=self.diff_panda.get_all_body_names())
        if len(shp) == 3:
            H = H.reshape(shp[0], shp[1], *H.shape[1:])
        return link_pos_from_link_tensor(H)

    def get_EE_pose(self, q):
        return self.diff_panda.compute_forward_kinematics_all_links(q, link_list=[self.link_name_ee])

    def get_EE_position(self, q):
        return link_pos_from_link_tensor(self.get_EE_pose(q))

    def get_EE_orientation(self, q, rotation_matrix=True):
        if not rotation_matrix:
            raise NotImplementedError("quaternion output is outside the hot path")
        return link_rot_from_link_tensor(self.get_EE_pose(q))
