"""RobotBase: configuration-space limits + collision model bookkeeping (reference:
robots/robot_base.py:13-182).  The collision model is data for the cost descriptor: which bodies carry
object-collision points and their radii, and which body pairs are checked against each other."""
from __future__ import annotations

import abc
import itertools
from abc import ABC
from math import ceil

import torch

from ..torch_planning_objectives.fields.distance_fields import CollisionSelfField
from ..torch_utils.torch_utils import to_numpy, to_torch


class RobotBase(ABC):
    def __init__(self, name="RobotBase", q_limits=None, grasped_object=None,
                 margin_for_grasped_object_collision_checking=0.001,
                 link_names_for_object_collision_checking=None, link_margins_for_object_collision_checking=None,
                 link_idxs_for_object_collision_checking=None, link_names_for_self_collision_checking=None,
                 link_names_pairs_for_self_collision_checking=None, link_idxs_for_self_collision_checking=None,
                 self_collision_margin_robot=0.001, link_names_for_self_collision_checking_with_grasped_object=None,
                 self_collision_margin_grasped_object=0.05, num_interpolated_points_for_self_collision_checking=1,
                 num_interpolated_points_for_object_collision_checking=1, dt=1.0, tensor_args=None,
                 grasped_object_point_idxs=None, **kwargs):
        """grasped_object_point_idxs: body indices of the grasped object's collision points (the subclass attaches
        them to its kinematic tree, DifferentiableTree.attach_points).  They take the place of the point block the
        reference concatenates after the link positions (robot_panda.py:155-168): same columns of `link_pos`,
        same margins (robot_base.py:75-79), same extra self-collision pairs (robot_base.py:118-127).  The
        reference's own field code mis-indexes that block and raises IndexError for any grasped object
        (distance_fields.py:137-154 selects the link columns twice), so what is implemented here is the behaviour
        robot_base.py sets up, checked against the reference's SDF primitives in tests/golden."""
        if grasped_object is not None and grasped_object_point_idxs is None:
            raise NotImplementedError("this robot does not attach grasped-object points to its kinematic tree")
        self.name = name
        self.tensor_args = tensor_args
        self.dt = dt
        assert q_limits is not None, "q_limits cannot be None"
        self.q_limits = q_limits
        self.q_min, self.q_max = q_limits[0], q_limits[1]
        self.q_min_np, self.q_max_np = to_numpy(self.q_min), to_numpy(self.q_max)
        self.q_distribution = torch.distributions.uniform.Uniform(self.q_min, self.q_max)
        self.q_dim = len(self.q_min)
        self.grasped_object = grasped_object
        n_grasped = 0 if grasped_object is None else grasped_object.n_base_points_for_collision
        grasped_idxs = [] if grasped_object is None else [int(i) for i in grasped_object_point_idxs]
        assert len(grasped_idxs) == n_grasped
        self.grasped_object_point_idxs = grasped_idxs
        self.margin_for_grasped_object_collision_checking = margin_for_grasped_object_collision_checking

        # ---- object-collision points: one per listed link (points_per_link == 1 in every shipped robot)
        n_links_obj = len(link_names_for_object_collision_checking)
        assert num_interpolated_points_for_object_collision_checking >= n_links_obj
        per_link = ceil(num_interpolated_points_for_object_collision_checking / n_links_obj)
        self.points_per_link_object_collision_checking = per_link
        self.num_interpolated_points_for_object_collision_checking = per_link * n_links_obj
        self.self_collision_margin_robot = self_collision_margin_robot
        self.link_names_for_object_collision_checking = link_names_for_object_collision_checking
        self.n_links_for_object_collision_checking = n_links_obj
        self.link_margins_for_object_collision_checking = link_margins_for_object_collision_checking
        self.link_margins_for_object_collision_checking_robot_tensor = torch.tensor(
            link_margins_for_object_collision_checking, **tensor_args).repeat_interleave(per_link)
        self.link_margins_for_object_collision_checking_tensor = self.link_margins_for_object_collision_checking_robot_tensor
        if n_grasped:
            self.link_margins_for_object_collision_checking_tensor = torch.cat(
                (self.link_margins_for_object_collision_checking_tensor,
                 torch.ones(n_grasped, **tensor_args) * margin_for_grasped_object_collision_checking))
        self.link_idxs_for_object_collision_checking = link_idxs_for_object_collision_checking
        # columns of fk_map_collision() the object / workspace fields read: the links, then the grasped points
        self.point_idxs_for_object_collision_checking = list(link_idxs_for_object_collision_checking) + grasped_idxs

        # ---- self-collision pairs (robot_base.py:87-141)
        if link_names_for_self_collision_checking is None:
            self.df_collision_self = None
            return
        n_links_self = len(link_names_for_self_collision_checking)
        assert num_interpolated_points_for_self_collision_checking >= n_links_self
        p = ceil(num_interpolated_points_for_self_collision_checking / n_links_self)
        self.points_per_link_self_collision_checking = p
        self.link_names_for_self_collision_checking = link_names_for_self_collision_checking
        self.link_names_pairs_for_self_collision_checking = link_names_pairs_for_self_collision_checking
        self.link_idxs_for_self_collision_checking = link_idxs_for_self_collision_checking
        self.link_names_for_self_collision_checking_with_grasped_object = link_names_for_self_collision_checking_with_grasped_object
        self.self_collision_margin_grasped_object = self_collision_margin_grasped_object
        pair_idxs = []
        for i, link_1 in enumerate(link_names_for_self_collision_checking):
            for link_2 in link_names_pairs_for_self_collision_checking.get(link_1, []):
                j = link_names_for_self_collision_checking.index(link_2)
                pair_idxs.extend((i * p + m, j * p + n) for m, n in itertools.product(range(p), range(p)))
        margin_list = [self_collision_margin_robot] * len(pair_idxs)
        rows = list(link_idxs_for_self_collision_checking)
        if n_grasped:
            last_row = n_links_self * p
            for link_1 in link_names_for_self_collision_checking_with_grasped_object:
                j = link_names_for_self_collision_checking.index(link_1)
                extra = [(last_row + m, j * p + n) for m, n in itertools.product(range(n_grasped), range(p))]
                pair_idxs.extend(extra)
                margin_list.extend([self_collision_margin_grasped_object] * len(extra))
            rows = rows + grasped_idxs
        margins = to_torch(margin_list, **tensor_args)
        self.df_collision_self = CollisionSelfField(
            self, link_idxs_for_collision_checking=rows,
            idxs_links_distance_matrix=pair_idxs, num_interpolated_points=p * n_links_self,
            cutoff_margin=margins, tensor_args=tensor_args)

    def random_q(self, n_samples=10):
        return self.q_distribution.sample((n_samples,))

    def get_position(self, x):
        return x[..., :self.q_dim]

    def get_velocity(self, x):
        return x[..., self.q_dim:2 * self.q_dim]

    def get_acceleration(self, x):
        return x[..., 2 * self.q_dim:3 * self.q_dim]

    def distance_q(self, q1, q2):
        return torch.linalg.norm(q1 - q2, dim=-1)

    def fk_map_collision(self, q, **kwargs):
        if q.ndim == 1:
            q = q.unsqueeze(0)
        return self.fk_map_collision_impl(q, **kwargs)

    @abc.abstractmethod
    def fk_map_collision_impl(self, q, **kwargs):
        raise NotImplementedError
