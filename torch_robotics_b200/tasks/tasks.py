"""PlanningTask: composes FK and the collision fields into a per-waypoint cost / collision flag
(reference: tasks/tasks.py:22-232).

`compute_collision_cost` and `compute_collision` keep their signatures; underneath, ONE fused kernel
launch evaluates FK of all links, the self / object / workspace terms and - for the cost - the analytic
gradient w.r.t. q (trk_objective), instead of FK + three field passes + autograd.  The individual
fields remain available through `get_collision_fields()` and give identical numbers.
`compute_objective` additionally folds an EESE3DistanceField into the same launch (BASELINE config 3).
"""
from __future__ import annotations

from abc import ABC

import torch

import sys

from .. import ops
from .._lib import DeviceCost
from ..descriptor import cost_in_translated_frame
from ..descriptor import build_cost_descriptor
from ..torch_planning_objectives.fields.distance_fields import (
    CollisionObjectDistanceField, CollisionWorkspaceBoundariesDistanceField, _tree_of)
from ..trajectory.utils import interpolate_traj_via_points


class Task(ABC):
    def __init__(self, env=None, robot=None, tensor_args=None, **kwargs):
        self.env = env
        self.robot = robot
        self.tensor_args = tensor_args


class PlanningTask(Task):
    def __init__(self, ws_limits=None, use_occupancy_map=False, cell_size=0.01, obstacle_cutoff_margin=0.01, **kwargs):
        super().__init__(**kwargs)
        if use_occupancy_map:
            raise NotImplementedError("occupancy maps are outside the hot path (the reference raises too, tasks.py:160)")
        self.ws_limits = self.env.limits if ws_limits is None else ws_limits
        self.ws_min, self.ws_max = self.ws_limits[0], self.ws_limits[1]
        self.use_occupancy_map = False
        self.obstacle_cutoff_margin = obstacle_cutoff_margin
        r = self.robot
        common = dict(link_idxs_for_collision_checking=r.point_idxs_for_object_collision_checking,
                      num_interpolated_points=r.num_interpolated_points_for_object_collision_checking,
                      link_margins_for_object_collision_checking_tensor=r.link_margins_for_object_collision_checking_tensor,
                      cutoff_margin=obstacle_cutoff_margin, tensor_args=self.tensor_args)
        self.df_collision_self = r.df_collision_self
        self.df_collision_objects = CollisionObjectDistanceField(r, df_obj_list_fn=self.env.get_df_obj_list, **common)
        self.df_collision_ws_boundaries = CollisionWorkspaceBoundariesDistanceField(
            r, ws_bounds_min=self.ws_min, ws_bounds_max=self.ws_max, **common)
        if self.env.obj_extra_list is not None:
            from functools import partial
            self.df_collision_extra_objects = CollisionObjectDistanceField(
                r, df_obj_list_fn=partial(self.env.get_df_obj_list, return_extra_objects_only=True), **common)
            self._collision_fields_extra_objects = [self.df_collision_extra_objects]
        else:
            self._collision_fields_extra_objects = []
        self._collision_fields = [self.df_collision_self, self.df_collision_objects, self.df_collision_ws_boundaries]
        self._fused = {}

    def get_collision_fields(self):
        return self._collision_fields

    def get_collision_fields_extra_objects(self):
        return self._collision_fields_extra_objects

    def distance_q(self, q1, q2):
        return self.robot.distance_q(q1, q2)

    # ---------------------------------------------------------------------------------- fused tables
    def cost_descriptor(self, ee_field=None, ee_link_name=None):
        """CostDescriptor equivalent to the three collision fields (+ optional EE field)."""
        r = self.robot
        tree = _tree_of(r)
        f_self, f_obj = self.df_collision_self, self.df_collision_objects
        pairs, self_margin = [], []
        if f_self is not None:
            links = list(f_self.link_idxs_for_collision_checking)
            pairs = [(links[a], links[b]) for a, b in f_self.idxs_links_distance_matrix]
            self_margin = (torch.as_tensor(f_self.cutoff_margin, dtype=torch.float32).reshape(-1).cpu() + 0.0).numpy()
        ee = {}
        if ee_field is not None:
            name = ee_link_name or getattr(r, "link_name_ee", None) or tree.get_link_names()[-1]
            ee = dict(ee_link=tree._name_to_idx_map[name], ee_target=torch.as_tensor(ee_field.target_H).reshape(4, 4),
                      ee_w_pos=ee_field.w_pos, ee_w_rot=ee_field.w_rot, ee_square=ee_field.square)
        return build_cost_descriptor(
            pt_link=r.point_idxs_for_object_collision_checking,
            pt_radius=r.link_margins_for_object_collision_checking_tensor.detach().float().cpu(),
            cutoff_obj=self.obstacle_cutoff_margin, objects=[o.as_sphere_table() for o in f_obj._objects()],
            ws_min=self.ws_min, ws_max=self.ws_max, self_pairs=pairs, self_margin=self_margin,
            clamp_self=bool(f_self.clamp_sdf) if f_self is not None else False,
            clamp_obj=f_obj.clamp_sdf, clamp_ws=self.df_collision_ws_boundaries.clamp_sdf, **ee)

    def _fused_cost(self, device, ee_field=None, ee_link_name=None) -> DeviceCost:
        tree = _tree_of(self.robot)
        sig, refs = self.df_collision_objects._signature()
        settings, srefs = [], []
        for f in (self.df_collision_self, self.df_collision_objects, self.df_collision_ws_boundaries):
            if f is not None:
                k, r = f._settings_key()
                settings.append(k)
                srefs += list(r)
        # keyed on content identity AND on the tree's model generation (update_base_pose / attach_* rebuild the device
        # model); the entry keeps every object whose id is part of the key alive, so ids cannot be recycled under it
        key = (torch.device(device).index, getattr(tree, "_model_generation", 0), sig, tuple(settings), id(ee_field), ee_link_name)
        # A base pose that is a pure translation (update_base_pose): the objective is evaluated in the robot's base frame - the
        # tree without the base pose keeps the unrolled kernels - on the scene shifted by -t (cost_in_translated_frame); costs,
        # gradients, occupancy and arg-min are frame-independent.  A rotated base stays on the runtime-tree kernels.
        shift = tree.base_translation() if hasattr(tree, "base_translation") else None
        entry = self._fused.get(key)
        if entry is None:
            if len(self._fused) > 8:
                self._fused.clear()
            desc = self.cost_descriptor(ee_field, ee_link_name)
            if shift is not None:
                model, desc = tree.device_model_in_base_frame(device), cost_in_translated_frame(desc, shift)
            else:
                model = tree.device_model(device)
            entry = (DeviceCost(model, desc), refs, tuple(srefs), ee_field)
            self._fused[key] = entry
        dc = entry[0]
        if ee_field is not None:
            tgt = torch.as_tensor(ee_field.target_H).reshape(4, 4).detach().float().cpu().numpy().copy()
            if shift is not None:
                tgt[:3, 3] -= shift
            if (tgt != dc.cost.ee_target).any():
                dc.set_ee_target(tgt)
        return dc

    @staticmethod
    def _as_bhd(q):
        if q.ndim == 1:
            return q.reshape(1, 1, -1), (1, 1)       # the reference's fields always answer [batch, horizon] (distance_fields.py:26-58)
        if q.ndim == 2:
            return q.unsqueeze(1), (q.shape[0], 1)
        if q.ndim == 3:
            return q, (q.shape[0], q.shape[1])
        raise NotImplementedError

    # ---------------------------------------------------------------------------------- public API
    def compute_collision(self, x, **kwargs):
        return self._compute_collision_or_cost(self.robot.get_position(x), field_type="occupancy", **kwargs)

    def compute_collision_cost(self, x, **kwargs):
        return self._compute_collision_or_cost(self.robot.get_position(x), field_type="sdf", **kwargs)

    def _compute_collision_or_cost(self, q, field_type="occupancy", **kwargs):
        """q [..., q_dim] with up to (batch, horizon) leading dims -> [batch, horizon] cost (fp32) or
        collision flag (bool).  Reference: tasks.py:139-232."""
        q3, out_shape = self._as_bhd(q)
        flat = q3.reshape(-1, q3.shape[-1])
        dc = self._fused_cost(flat.device)
        if field_type == "sdf":
            return ops.objective(dc, flat).reshape(out_shape)
        if field_type == "occupancy":
            r = dc.objective(flat.detach(), want_grad=False, want_terms=False, want_occupancy=True,
                             margin_override=kwargs.get("margin", None))
            return (r["occupancy"] != 0).reshape(out_shape)
        raise NotImplementedError(field_type)

    def compute_objective(self, x, ee_field, ee_link_name=None):
        """Collision cost + EESE3DistanceField cost of the EE link in one fused launch -> [batch, horizon]."""
        q3, out_shape = self._as_bhd(self.robot.get_position(x))
        flat = q3.reshape(-1, q3.shape[-1])
        return ops.objective(self._fused_cost(flat.device, ee_field, ee_link_name), flat).reshape(out_shape)

    def compute_cost_and_grad(self, x, ee_field=None, ee_link_name=None):
        """(cost [batch, horizon], d cost.sum() / d q [batch, horizon, D]) from ONE kernel launch, without autograd:
        the call for planners that consume the gradient directly (GPMP / CHOMP style updates).  Equal to
        `c = compute_collision_cost(q)` (or `compute_objective`) followed by `c.sum().backward()`, minus the autograd
        bookkeeping (~100 us per call, which at 4096 x 64 configurations is three times the kernel itself)."""
        q3, out_shape = self._as_bhd(self.robot.get_position(x))
        flat = q3.detach().reshape(-1, q3.shape[-1])
        r = self._fused_cost(flat.device, ee_field, ee_link_name).objective(flat, want_grad=True, want_terms=False)
        return r["total"].reshape(out_shape), r["grad_q"].reshape(*out_shape, q3.shape[-1])

    # ---------------------------------------------------------------------------------- sampling
    def sample_q(self, without_collision=True, **kwargs):
        return self.random_coll_free_q(**kwargs) if without_collision else self.robot.random_q(**kwargs)

    def random_coll_free_q(self, n_samples=1, max_samples=1000, max_tries=1000):
        """Rejection sampling of collision-free configurations (reference: tasks.py:104-129)."""
        samples = torch.zeros((n_samples, self.robot.q_dim), **self.tensor_args)
        filled = 0
        for _ in range(max_tries):
            qs = self.robot.random_q(max_samples)
            free = torch.nonzero(~self.compute_collision(qs).reshape(-1)).reshape(-1)
            if free.numel() == 0:
                continue
            pick = free[torch.randperm(free.numel(), device=free.device)[:n_samples]]
            take = min(pick.numel(), n_samples - filled)
            samples[filled:filled + take] = qs[pick[:take]]
            filled += take
            if filled >= n_samples:
                return samples.squeeze()
        sys.exit("Could not find a collision free configuration")

    # ---------------------------------------------------------------------------------- trajectories
    def get_trajs_collision_and_free(self, trajs, return_indices=False, num_interpolation=5):
        """Split a batch of trajectories [B, H, D] (or [N, B, H, D]) into colliding and collision-free ones
        (reference: tasks.py:234-308).  Collisions are tested with zero margin on the points interpolated
        between the waypoints; a collision-free trajectory that leaves the joint limits counts as colliding.
        Index tensors are int64 [n, 1] ([n, 2] for 4-D input), in the reference's order: trajectories with a
        colliding waypoint first, then the limit violators."""
        assert trajs.ndim in (3, 4)
        four_d = trajs.ndim == 4
        flat = trajs.reshape(-1, trajs.shape[-2], trajs.shape[-1]) if four_d else trajs
        waypoint_coll = self.compute_collision(interpolate_traj_via_points(flat, num_interpolation=num_interpolation), margin=0.)
        if four_d:
            waypoint_coll = waypoint_coll.reshape(trajs.shape[0], trajs.shape[1], -1)
        free_idx = torch.argwhere(~waypoint_coll.any(dim=-1))
        coll_idx = torch.argwhere(waypoint_coll.any(dim=-1))
        if free_idx.numel() > 0:
            sel = trajs[tuple(free_idx.t())]
            pos = self.robot.get_position(sel)
            inside = ((pos >= self.robot.q_min) & (pos <= self.robot.q_max)).all(dim=-1).all(dim=-1)
            if not bool(inside.any()):
                coll_idx, free_idx = free_idx.clone(), free_idx[:0]
            else:
                coll_idx = torch.cat((coll_idx, free_idx[~inside]))
                free_idx = free_idx[inside]
        trajs_free = trajs[tuple(free_idx.t())] if free_idx.numel() else None
        trajs_coll = trajs[tuple(coll_idx.t())] if coll_idx.numel() else None
        if four_d:
            trajs_free = None if trajs_free is None else trajs_free
            trajs_coll = None if trajs_coll is None else trajs_coll
        if return_indices:
            return trajs_coll, coll_idx, trajs_free, free_idx, waypoint_coll
        return trajs_coll, trajs_free

    def compute_fraction_free_trajs(self, trajs, **kwargs):
        _, coll_idx, _, free_idx, _ = self.get_trajs_collision_and_free(trajs, return_indices=True)
        n_free, n_coll = free_idx.nelement(), coll_idx.nelement()
        return n_free / (n_free + n_coll)

    def compute_collision_intensity_trajs(self, trajs, **kwargs):
        wp = self.get_trajs_collision_and_free(trajs, return_indices=True)[4]
        return torch.count_nonzero(wp) / wp.nelement()

    def compute_success_free_trajs(self, trajs, **kwargs):
        free = self.get_trajs_collision_and_free(trajs)[1]
        return 1 if free is not None and free.nelement() >= 1 else 0
