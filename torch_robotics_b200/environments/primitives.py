"""Obstacle primitives the fused objective can ingest (reference: environments/primitives.py:20-112,
348-405).  These are host-side descriptions: `compute_signed_distance` evaluates them with eager torch
for API parity, while PlanningTask / the DistanceField classes flatten them into the CostDescriptor
tables read by the CUDA kernels (spheres today; boxes are SURVEY.md section 8 row f2)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List

import torch

from ..torch_kinematics_tree.geometrics.quaternion import q_to_rotation_matrix
from ..torch_utils.torch_utils import DEFAULT_TENSOR_ARGS, to_torch


class PrimitiveShapeField(ABC):
    def __init__(self, dim=3, tensor_args=None):
        self.dim = dim
        self.tensor_args = DEFAULT_TENSOR_ARGS if tensor_args is None else tensor_args

    def compute_signed_distance(self, x):
        return self.compute_signed_distance_impl(x)

    @abstractmethod
    def compute_signed_distance_impl(self, x):
        raise NotImplementedError

    def zero_grad(self):
        pass


class MultiSphereField(PrimitiveShapeField):
    def __init__(self, centers, radii, tensor_args=None):
        super().__init__(dim=centers.shape[-1], tensor_args=tensor_args)
        self.centers = to_torch(centers, **self.tensor_args)
        self.radii = to_torch(radii, **self.tensor_args)

    def __repr__(self):
        return f"MultiSphereField(centers={self.centers}, radii={self.radii})"

    def compute_signed_distance_impl(self, x):
        d = torch.linalg.norm(x.unsqueeze(-2) - self.centers, dim=-1) - self.radii
        return d.min(dim=-1)[0]

    def zero_grad(self):
        self.centers.grad = None
        self.radii.grad = None


class MultiSharpBoxField(PrimitiveShapeField):
    """Axis-aligned boxes, Chebyshev-style signed distance (reference: primitives.py:198-223)."""
    ROUNDED = False

    def __init__(self, centers, sizes, tensor_args=None):
        super().__init__(dim=centers.shape[-1], tensor_args=tensor_args)
        self.centers = to_torch(centers, **self.tensor_args)
        self.sizes = to_torch(sizes, **self.tensor_args)
        self.half_sizes = self.sizes / 2

    def __repr__(self):
        return f"{type(self).__name__}(centers={self.centers}, sizes={self.sizes})"

    def compute_signed_distance_impl(self, x):
        d = torch.abs(x.unsqueeze(-2) - self.centers) - self.half_sizes
        return d.max(dim=-1)[0].min(dim=-1)[0]


class MultiBoxField(MultiSharpBoxField):
    """Rounded boxes, radius 0.15 x the smallest side (reference: primitives.py:309-334)."""
    ROUNDED = True

    def __init__(self, centers, sizes, tensor_args=None):
        super().__init__(centers, sizes, tensor_args=tensor_args)
        self.radius = torch.min(self.sizes, dim=-1)[0] * 0.15

    def compute_signed_distance_impl(self, x):
        q = torch.abs(x.unsqueeze(-2) - self.centers) - self.half_sizes + self.radius.unsqueeze(-1)
        mq = torch.amax(q, dim=-1)
        sdf = torch.minimum(mq, torch.zeros_like(mq)) + torch.linalg.norm(torch.relu(q), dim=-1) - self.radius
        return sdf.min(dim=-1)[0]


MultiRoundedBoxField = MultiBoxField


class ObjectField(PrimitiveShapeField):
    """A posed group of primitive fields (pose = pos + wxyz quaternion in the world frame)."""

    def __init__(self, primitive_fields, name="object", pos=None, ori=None, reference_frame="base"):
        assert isinstance(primitive_fields, List) and len(primitive_fields) > 0
        super().__init__(dim=primitive_fields[0].dim, tensor_args=primitive_fields[0].tensor_args)
        self.name = name
        self.fields = primitive_fields
        assert (pos is None and ori is None) or (pos.nelement() == 3 and ori.nelement() == 4)
        self.pos = torch.zeros(3, **self.tensor_args) if pos is None else pos
        self.ori = torch.tensor([1, 0, 0, 0], **self.tensor_args) if ori is None else ori
        self.reference_frame = reference_frame
        self.grid_map_sdf = None

    def __repr__(self):
        return f"ObjectField(fields={self.fields})"

    def set_position_orientation(self, pos=None, ori=None):
        self._pose_revision = getattr(self, "_pose_revision", 0) + 1      # cache keys of the fused cost tables see the move
        if pos is not None:
            assert len(pos) == 3
            self.pos = to_torch(pos, **self.tensor_args)
        if ori is not None:
            assert len(ori) == 4, "quaternion wxyz"
            self.ori = to_torch(ori, **self.tensor_args)

    def compute_signed_distance_impl(self, x):
        R = q_to_rotation_matrix(self.ori).squeeze(0)
        x_obj = (x - self.pos) @ R          # R^T (x - pos), row-vector form
        sdf = torch.stack([f.compute_signed_distance_impl(x_obj) for f in self.fields], dim=-1)
        return sdf.min(dim=-1)[0]

    def zero_grad(self):
        for f in self.fields:
            f.zero_grad()

    def as_sphere_table(self):
        """(centers [K,3], radii [K], pos [3], ori_wxyz [4], boxes) for the cost descriptor: sphere fields are
        concatenated; box fields are listed as (centers, sizes, rounded)."""
        cs, rs, boxes = [torch.zeros(0, 3)], [torch.zeros(0)], []
        for f in self.fields:
            if isinstance(f, MultiSphereField):
                cs.append(f.centers.detach().float().cpu().reshape(-1, 3))
                rs.append(f.radii.detach().float().cpu().reshape(-1))
            elif isinstance(f, MultiSharpBoxField):
                boxes.append((f.centers.detach().float().cpu().numpy(), f.sizes.detach().float().cpu().numpy(), f.ROUNDED))
            else:
                raise NotImplementedError(f"{type(f).__name__} obstacles are not ingested by the fused kernel")
        return torch.cat(cs).numpy(), torch.cat(rs).numpy(), self.pos.detach().float().cpu().numpy(), \
            self.ori.detach().float().cpu().numpy(), boxes
