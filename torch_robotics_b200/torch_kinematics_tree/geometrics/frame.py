"""SE(3) frame value object with the reference's surface (geometrics/frame.py:12-129): what FK returns
with return_dict=True, and what callers use to build targets."""
from __future__ import annotations

from typing import Optional

import torch

from .quaternion import q_to_rotation_matrix


class Frame:
    def __init__(self, rot: Optional[torch.Tensor] = None, trans: Optional[torch.Tensor] = None,
                 pose: Optional[torch.Tensor] = None, device="cpu"):
        self.device = device
        self._rot = torch.eye(3, device=device).unsqueeze(0) if rot is None else rot.to(device)
        if self._rot.ndim == 2:
            self._rot = self._rot.unsqueeze(0)
        self._trans = torch.zeros(1, 3, device=device) if trans is None else trans.to(device)
        if self._trans.ndim == 1:
            self._trans = self._trans.unsqueeze(0)
        if pose is not None:
            self.set_pose(pose)
        assert self._trans.shape[0] == self._rot.shape[0]
        self.batch_size = self._trans.shape[0]

    def set_pose(self, pose: torch.Tensor) -> None:
        """pose rows: x, y, z, qw, qx, qy, qz"""
        pose = pose.unsqueeze(0) if pose.ndim == 1 else pose
        self._trans = pose[:, :3].clone().to(self.device)
        self._rot = q_to_rotation_matrix(pose[:, 3:]).to(self.device)

    def set_translation(self, t):
        self._trans = t.to(self.device)

    def set_rotation(self, rot):
        self._rot = rot.to(self.device)

    def inverse(self) -> "Frame":
        Rt = self._rot.transpose(-2, -1)
        return Frame(Rt, -(Rt @ self._trans.unsqueeze(-1)).squeeze(-1), device=self.device)

    def multiply_transform(self, other: "Frame") -> "Frame":
        R = self._rot @ other.rotation
        t = (self._rot @ other.translation.float().unsqueeze(-1)).squeeze(-1) + self._trans
        return Frame(R, t, device=self.device)

    def get_transform_matrix(self) -> torch.Tensor:
        H = torch.eye(4, device=self._rot.device, dtype=self._rot.dtype).repeat(self.batch_size, 1, 1)
        H[:, :3, :3] = self._rot
        H[:, :3, 3] = self._trans
        return H

    def transform_point(self, point: torch.Tensor) -> torch.Tensor:
        """point [P,3] -> [B,P,3] in the frame's parent coordinates."""
        return (self._rot @ point.unsqueeze(0).transpose(-1, -2)).transpose(-1, -2) + self._trans.unsqueeze(-2)

    def get_euler(self):
        R = self._rot
        return torch.atan2(R[:, 2, 1], R[:, 2, 2]), torch.asin(-R[:, 2, 0]), torch.atan2(R[:, 1, 0], R[:, 0, 0])

    @property
    def rotation(self) -> torch.Tensor:
        return self._rot

    @property
    def translation(self) -> torch.Tensor:
        return self._trans
