"""Axis rotations with the reference's names (geometrics/spatial_vector.py:8-47).  Host-side helpers
for building targets and base poses; the kernels have their own per-joint rotation."""
import torch


def _axis_rot(k, angle):
    angle = torch.as_tensor(angle)
    angle = angle.reshape(-1).to(torch.get_default_dtype()) if not angle.is_floating_point() else angle.reshape(-1)
    c, s = torch.cos(angle), torch.sin(angle)
    R = torch.zeros((angle.shape[0], 3, 3), device=angle.device, dtype=angle.dtype)
    i, j = [(1, 2), (2, 0), (0, 1)][k]
    R[:, k, k] = 1.0
    R[:, i, i] = c
    R[:, j, j] = c
    R[:, i, j] = -s
    R[:, j, i] = s
    return R


def x_rot(angle):
    return _axis_rot(0, angle)


def y_rot(angle):
    return _axis_rot(1, angle)


def z_rot(angle):
    return _axis_rot(2, angle)
