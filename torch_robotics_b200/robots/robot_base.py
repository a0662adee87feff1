"""RobotBase: configuration-space limits + collision model bookkeeping (reference:
robots/robot_base.py:13-182).  The collision model is data for the cost descriptor: which bodies carry
object-collision points and their radii, and which body pairs are checked against each other."""
from __future__ import annotations

import abc
from abc import ABC
from math import ceil

import torch

from ..torch_planning_objectives.fields.distance_fields import CollisionSelfField
from ..torch_utils.torch_utils import to_numpy, to_torch


def _points_per_link(requested: int, n_links: int) -> int:
    """Interpolated points per link: the requested total rounded up to a multiple of the link count (robot_base.py:57-64)."""
    assert requested >= n_links
    return ceil(requested / n_links)


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
        assert q_limits is not None, "q_limits cannot be None"
        self.name, self.tensor_args, self.dt = name, tensor_args, dt
        self._init_configuration_space(q_limits)
        self.grasped_object = grasped_object
        self.grasped_object_point_idxs = [] if grasped_object is None else [int(i) for i in grasped_object_point_idxs]
        assert len(self.grasped_object_point_idxs) == (0 if grasped_object is None else grasped_object.n_base_points_for_collision)
        self.margin_for_grasped_object_collision_checking = margin_for_grasped_object_collision_checking
        self.self_collision_margin_robot = self_collision_margin_robot
        self._init_object_collision_model(link_names_for_object_collision_checking, link_margins_for_object_collision_checking,
                                          link_idxs_for_object_collision_checking,
                                          num_interpolated_points_for_object_collision_checking)
        self.df_collision_self = None
        if link_names_for_self_collision_checking is not None:
            self._init_self_collision_model(link_names_for_self_collision_checking, link_names_pairs_for_self_collision_checking,
                                            link_idxs_for_self_collision_checking,
                                            link_names_for_self_collision_checking_with_grasped_object,
                                            self_collision_margin_grasped_object,
                                            num_interpolated_points_for_self_collision_checking)

    # ---------------------------------------------------------------------------------- construction helpers
    def _init_configuration_space(self, q_limits):
        self.q_limits = q_limits
        self.q_min, self.q_max = q_limits[0], q_limits[1]
        self.q_min_np, self.q_max_np = to_numpy(self.q_min), to_numpy(self.q_max)
        self.q_dim = len(self.q_min)
        self.q_distribution = torch.distributions.uniform.Uniform(self.q_min, self.q_max)

    def _init_object_collision_model(self, names, radii, body_idxs, n_points_requested):
        """Collision points against obstacles / workspace: one sphere per listed link origin (every shipped robot has
        exactly one point per link), then the grasped object's points with their own small margin."""
        per_link = _points_per_link(n_points_requested, len(names))
        self.points_per_link_object_collision_checking = per_link
        self.num_interpolated_points_for_object_collision_checking = per_link * len(names)
        self.link_names_for_object_collision_checking = names
        self.n_links_for_object_collision_checking = len(names)
        self.link_margins_for_object_collision_checking = radii
        self.link_idxs_for_object_collision_checking = body_idxs
        robot_radii = torch.tensor(radii, **self.tensor_args).repeat_interleave(per_link)
        self.link_margins_for_object_collision_checking_robot_tensor = robot_radii
        n_grasped = len(self.grasped_object_point_idxs)
        grasp_radii = torch.full((n_grasped,), 1.0, **self.tensor_args) * self.margin_for_grasped_object_collision_checking
        self.link_margins_for_object_collision_checking_tensor = torch.cat((robot_radii, grasp_radii)) if n_grasped else robot_radii
        # columns of fk_map_collision() the object / workspace fields read: the links, then the grasped points
        self.point_idxs_for_object_collision_checking = list(body_idxs) + self.grasped_object_point_idxs

    def _init_self_collision_model(self, names, pairs_by_name, body_idxs, names_vs_grasped, margin_grasped, n_points_requested):
        """Self-collision pairs as (row, row) indices into the list [links..., grasped points...] (robot_base.py:87-141),
        with one margin per pair."""
        p = _points_per_link(n_points_requested, len(names))
        self.points_per_link_self_collision_checking = p
        self.link_names_for_self_collision_checking = names
        self.link_names_pairs_for_self_collision_checking = pairs_by_name
        self.link_idxs_for_self_collision_checking = body_idxs
        self.link_names_for_self_collision_checking_with_grasped_object = names_vs_grasped
        self.self_collision_margin_grasped_object = margin_grasped
        row_of = {n: k for k, n in enumerate(names)}

        def block(rows_a, rows_b):
            return [(a, b) for a in rows_a for b in rows_b]

        pair_rows, pair_margins = [], []
        for first in names:
            for second in pairs_by_name.get(first, []):
                new = block(range(row_of[first] * p, (row_of[first] + 1) * p), range(row_of[second] * p, (row_of[second] + 1) * p))
                pair_rows += new
                pair_margins += [self.self_collision_margin_robot] * len(new)
        rows = list(body_idxs)
        n_grasped = len(self.grasped_object_point_idxs)
        if n_grasped:
            first_grasp_row = len(names) * p
            for link in names_vs_grasped:
                new = block(range(first_grasp_row, first_grasp_row + n_grasped), range(row_of[link] * p, (row_of[link] + 1) * p))
                pair_rows += new
                pair_margins += [margin_grasped] * len(new)
            rows += self.grasped_object_point_idxs
        self.df_collision_self = CollisionSelfField(
            self, link_idxs_for_collision_checking=rows, idxs_links_distance_matrix=pair_rows,
            num_interpolated_points=p * len(names), cutoff_margin=to_torch(pair_margins, **self.tensor_args),
            tensor_args=self.tensor_args)

    # ---------------------------------------------------------------------------------- state helpers
    def random_q(self, n_samples=10):
        return self.q_distribution.sample((n_samples,))

    def get_position(self, x):
        return self._state_block(x, 0)

    def _state_block(self, x, k):
        return x[..., k * self.q_dim:(k + 1) * self.q_dim]

    def get_velocity(self, x):
        """Velocity block of the state; a position-only trajectory is differentiated (central differences, self.dt)."""
        v = self._state_block(x, 1)
        if x.nelement() != 0 and v.nelement() == 0:
            from ..trajectory.utils import finite_difference_vector
            return finite_difference_vector(x, dt=self.dt, method="central")
        return v

    def get_acceleration(self, x):
        a = self._state_block(x, 2)
        if x.nelement() != 0 and a.nelement() == 0:
            from ..trajectory.utils import finite_difference_vector
            return finite_difference_vector(self.get_velocity(x), dt=self.dt, method="central")
        return a

    def distance_q(self, q1, q2):
        return torch.linalg.norm(q1 - q2, dim=-1)

    def fk_map_collision(self, q, **kwargs):
        if q.ndim == 1:
            q = q.unsqueeze(0)
        return self.fk_map_collision_impl(q, **kwargs)

    @abc.abstractmethod
    def fk_map_collision_impl(self, q, **kwargs):
        raise NotImplementedError
