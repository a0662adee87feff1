"""SE(3) frame value object with the reference's surface (geometrics/frame.py:12-129): what FK returns
with return_dict=True, and what callers use to build targets.  Batched: rotation [B,3,3], translation [B,3]."""
from __future__ import annotations

from typing import Optional

import torch

from .quaternion import q_to_rotation_matrix


def _batched(t: torch.Tensor, ndim: int) -> torch.Tensor:
    """Add the leading batch axis to an unbatched rotation (2-D) / translation (1-D)."""
    return t.unsqueeze(0) if t.ndim == ndim - 1 else t


class Frame:
    def __init__(self, rot: Optional[torch.Tensor] = None, trans: Optional[torch.Tensor] = None,
                 pose: Optional[torch.Tensor] = None, device="cpu"):
        self.device = device
        self._rot = _batched(torch.eye(3, device=device) if rot is None else rot.to(device), 3)
        self._trans = _batched(torch.zeros(3, device=device) if trans is None else trans.to(device), 2)
        if pose is not None:
            self.set_pose(pose)
        assert self._trans.shape[0] == self._rot.shape[0]
        self.batch_size = self._trans.shape[0]

    # ---- accessors
    @property
    def rotation(self) -> torch.Tensor:
        return self._rot

    @property
    def translation(self) -> torch.Tensor:
        return self._trans

    def set_rotation(self, rot):
        self._rot = rot.to(self.device)

    def set_translation(self, t):
        self._trans = t.to(self.device)

    def set_pose(self, pose: torch.Tensor) -> None:
        """pose rows: x, y, z, qw, qx, qy, qz"""
        pose = _batched(pose, 2)
        self._rot = q_to_rotation_matrix(pose[:, 3:]).to(self.device)
        self._trans = pose[:, :3].clone().to(self.device)

    # ---- group operations
    def multiply_transform(self, other: "Frame") -> "Frame":
        """self o other: rotation R_s R_o, translation R_s t_o + t_s (fp32, geometrics/utils.py:11-17)."""
        t = torch.einsum("bij,bj->bi", self._rot, other.translation.float()) + self._trans
        return Frame(self._rot @ other.rotation, t, device=self.device)

    def inverse(self) -> "Frame":
        Rt = self._rot.transpose(-2, -1)
        return Frame(Rt, -torch.einsum("bij,bj->bi", Rt, self._trans), device=self.device)

    def transform_point(self, point: torch.Tensor) -> torch.Tensor:
        """point [P,3] (or [3]) -> [B,P,3] in the frame's parent coordinates."""
        return torch.einsum("bij,pj->bpi", self._rot, point.reshape(-1, 3)) + self._trans[:, None, :]

    # ---- conversions
    def get_transform_matrix(self) -> torch.Tensor:
        H = torch.zeros(self.batch_size, 4, 4, device=self._rot.device, dtype=self._rot.dtype)
        H[:, :3, :3], H[:, :3, 3], H[:, 3, 3] = self._rot, self._trans, 1.0
        return H

    def get_euler(self):
        """(roll, pitch, yaw) of the Z-Y-X convention."""
        R = self._rot
        return torch.atan2(R[:, 2, 1], R[:, 2, 2]), torch.asin(-R[:, 2, 0]), torch.atan2(R[:, 1, 0], R[:, 0, 0])
