"""Cost fields on link positions / poses with the reference's class names and call signatures
(torch_planning_objectives/fields/distance_fields.py:12-359).

Each field owns a one-term CostDescriptor; `compute_cost` hands the caller's link-position tensor to the
cost-stage kernel (trk_cost_stage) through an autograd.Function, so costs stay differentiable w.r.t. the
FK output exactly as in the reference.  When all fields are evaluated together, PlanningTask uses the
fused FK+objective kernel instead (tasks/tasks.py).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
import torch

from ... import ops
from ..._lib import DeviceCost
from ...descriptor import build_cost_descriptor
from ...torch_kinematics_tree.geometrics.utils import SE3_distance


def _tree_of(robot):
    for attr in ("diff_model", "diff_panda"):
        tree = getattr(robot, attr, None)
        if tree is not None:
            return tree
    raise RuntimeError("the robot passed to a distance field must expose its DifferentiableTree "
                       "as `diff_model` (RobotPanda: `diff_panda`)")


class DistanceField(ABC):
    def __init__(self, tensor_args=None):
        self.tensor_args = tensor_args

    def compute_cost(self, q, link_pos, *args, **kwargs):
        """link_pos: [b,d] | [b,t,d] | [b,h,t,3] positions, or [b,h,t,4,4] poses -> cost [b,h]
        (reference: distance_fields.py:26-55)."""
        shp = link_pos.shape
        if len(shp) == 2:
            b, h = shp[0], 1
            link_pos = link_pos.unsqueeze(1)
        elif len(shp) == 3:
            b, h = shp[0], 1
        elif len(shp) == 4:
            b, h = shp[0], shp[1]
            link_pos = link_pos.reshape(b * h, shp[2], shp[3])
        elif len(shp) == 5:
            b, h = shp[0], shp[1]
            link_pos = link_pos.reshape(b * h, shp[2], shp[3], shp[4])
        else:
            raise NotImplementedError
        cost = self.compute_costs_impl(q, link_pos, *args, **kwargs)
        if cost.ndim == 1:
            cost = cost.reshape(b, h)
        return cost

    @abstractmethod
    def compute_costs_impl(self, *args, **kwargs):
        pass

    def zero_grad(self):
        pass


class EmbodimentDistanceFieldBase(DistanceField):
    """Shared plumbing of the three collision fields (distance_fields.py:82-177)."""

    TERM_BIT = 0

    def __init__(self, robot, link_idxs_for_collision_checking=None, num_interpolated_points=30, collision_margins=0.,
                 cutoff_margin=0.001, field_type="sdf", clamp_sdf=False, interpolate_link_pos=False, **kwargs):
        super().__init__(**kwargs)
        assert robot is not None, "You need to pass a robot instance to the embodiment distance fields"
        if interpolate_link_pos:
            raise NotImplementedError("interpolate_link_pos=True is outside the hot path (default False, :94)")
        self.robot = robot
        self.link_idxs_for_collision_checking = link_idxs_for_collision_checking
        self.num_interpolated_points = num_interpolated_points
        self.collision_margins = collision_margins
        self.cutoff_margin = cutoff_margin
        self.field_type = field_type
        self.clamp_sdf = clamp_sdf
        self.interpolate_link_pos = interpolate_link_pos
        self._device_costs = {}

    # -- subclasses describe their single term
    @abstractmethod
    def _descriptor(self):
        raise NotImplementedError

    def _signature(self):
        """(hashable key, objects the key's ids refer to).  The cache entry keeps the objects alive, so an id in the key
        cannot be recycled by CPython for a different tensor while the entry exists."""
        return None, ()

    def _settings_key(self):
        cm = self.cutoff_margin
        cm_key = (id(cm), cm._version) if torch.is_tensor(cm) else float(cm)
        return (bool(self.clamp_sdf), cm_key), ((cm,) if torch.is_tensor(cm) else ())

    def _device_cost(self, device) -> DeviceCost:
        tree = _tree_of(self.robot)
        sig, refs = self._signature()
        skey, srefs = self._settings_key()
        # the tree's generation changes with update_base_pose / attach_*: tables built for an older model are dropped
        key = (torch.device(device).index, getattr(tree, "_model_generation", 0), sig, skey)
        entry = self._device_costs.get(key)
        if entry is None:
            self._device_costs.clear()
            entry = (DeviceCost(tree.device_model(device), self._descriptor()), refs, srefs)
            self._device_costs[key] = entry
        return entry[0]

    def compute_costs_impl(self, q, link_pos, field_type=None, **kwargs):
        # grasped-object points are ordinary columns of link_pos here (bodies attached to the tree, robot_base.py), so
        # the reference's split / re-concatenation (distance_fields.py:137-152) has nothing to do
        field_type = self.field_type if field_type is None else field_type
        dc = self._device_cost(link_pos.device)
        if field_type == "sdf":
            return ops.cost_stage(dc, link_pos)
        if field_type == "occupancy":
            r = dc.cost_stage(link_pos.detach(), want_grad=False, want_occupancy=True,
                              margin_override=kwargs.get("margin", None))
            return (r["occupancy"] & self.TERM_BIT) != 0
        raise NotImplementedError(f"field_type {field_type} not implemented")

    def compute_distance(self, q, link_pos, **kwargs):
        raise NotImplementedError


class CollisionSelfField(EmbodimentDistanceFieldBase):
    """Pairwise link-origin distances against a per-pair margin (distance_fields.py:180-215)."""
    TERM_BIT = 1

    def __init__(self, *args, idxs_links_distance_matrix=None, **kwargs):
        super().__init__(*args, collision_margins=0., **kwargs)
        self.idxs_links_distance_matrix = idxs_links_distance_matrix
        self.idxs_links_distance_matrix_tuple = tuple(zip(*idxs_links_distance_matrix))

    def _descriptor(self):
        links = list(self.link_idxs_for_collision_checking)
        if len(links) < 2:
            raise NotImplementedError("self-collision with a single link (distance_fields.py:195-198)")
        pairs = [(links[a], links[b]) for a, b in self.idxs_links_distance_matrix]
        margin = torch.as_tensor(self.cutoff_margin, dtype=torch.float32).reshape(-1).cpu() + 0.0
        return build_cost_descriptor(self_pairs=pairs, self_margin=margin.numpy(), clamp_self=self.clamp_sdf,
                                     weights=(1.0, 0.0, 0.0, 0.0))


class CollisionObjectBase(EmbodimentDistanceFieldBase):
    def __init__(self, *args, link_margins_for_object_collision_checking_tensor=None, **kwargs):
        super().__init__(*args, collision_margins=link_margins_for_object_collision_checking_tensor, **kwargs)

    def _points(self):
        return list(self.link_idxs_for_collision_checking), torch.as_tensor(self.collision_margins).detach().float().cpu()


class CollisionObjectDistanceField(CollisionObjectBase):
    """Collision points against the environment's ObjectFields (distance_fields.py:298-316)."""
    TERM_BIT = 2

    def __init__(self, *args, df_obj_list_fn=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.df_obj_list_fn = df_obj_list_fn

    def _objects(self):
        return [] if self.df_obj_list_fn is None else list(self.df_obj_list_fn())

    def _signature(self):
        sig, refs = [], []
        for obj in self._objects():
            if not hasattr(obj, "fields"):      # GridMapSDF
                sig.append((id(obj), id(obj.grid)))
                refs += [obj, obj.grid]
                continue
            geo = []
            for f in obj.fields:
                shape = getattr(f, "radii", getattr(f, "sizes", None))
                geo.append((id(f.centers), f.centers._version, id(shape), shape._version))
                refs += [f.centers, shape]
            sig.append((id(obj), getattr(obj, "_pose_revision", 0), id(obj.pos), obj.pos._version, id(obj.ori), obj.ori._version, tuple(geo)))
            refs += [obj, obj.pos, obj.ori]
        return tuple(sig), tuple(refs)

    def _descriptor(self):
        links, radius = self._points()
        return build_cost_descriptor(pt_link=links, pt_radius=radius, cutoff_obj=self.cutoff_margin,
                                     objects=[o.as_sphere_table() for o in self._objects()], clamp_obj=self.clamp_sdf,
                                     weights=(0.0, 1.0, 0.0, 0.0))


class CollisionWorkspaceBoundariesDistanceField(CollisionObjectBase):
    """Collision points against the six workspace half-spaces (distance_fields.py:319-332)."""
    TERM_BIT = 4

    def __init__(self, *args, ws_bounds_min=None, ws_bounds_max=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.ws_min = ws_bounds_min
        self.ws_max = ws_bounds_max

    def _descriptor(self):
        links, radius = self._points()
        return build_cost_descriptor(pt_link=links, pt_radius=radius, cutoff_obj=self.cutoff_margin,
                                     ws_min=self.ws_min, ws_max=self.ws_max, clamp_ws=self.clamp_sdf,
                                     weights=(0.0, 0.0, 1.0, 0.0))


class EESE3DistanceField(DistanceField):
    """SE(3) distance of the last link of the tensor it is given to a target pose
    (distance_fields.py:335-359)."""

    def __init__(self, target_H, w_pos=1., w_rot=1., square=True, **kwargs):
        super().__init__(**kwargs)
        self.target_H = target_H
        self.square = square
        self.w_pos = w_pos
        self.w_rot = w_rot

    def update_target(self, target_H):
        self.target_H = target_H

    def compute_distance(self, link_tensor):
        return SE3_distance(link_tensor[..., -1, :, :], self.target_H, w_pos=self.w_pos, w_rot=self.w_rot)

    def compute_costs_impl(self, q, link_tensor, **kwargs):
        dist = self.compute_distance(link_tensor).squeeze()
        return torch.square(dist) if self.square else dist

    def zero_grad(self):
        raise NotImplementedError
