"""The two quaternion helpers the hot path's callers touch (reference: geometrics/quaternion.py:102-120)."""
import torch


def q_to_rotation_matrix(q):
    """wxyz quaternion(s) -> rotation matrix; not required to be unit length."""
    q = q.unsqueeze(0) if q.ndim == 1 else q
    w, x, y, z = torch.unbind(q, -1)
    k = 2.0 / (q ** 2).sum(-1)
    rows = (1 - k * (y * y + z * z), k * (x * y - z * w), k * (x * z + y * w),
            k * (x * y + z * w), 1 - k * (x * x + z * z), k * (y * z - x * w),
            k * (x * z - y * w), k * (y * z + x * w), 1 - k * (x * x + y * y))
    return torch.stack(rows, dim=-1).reshape(q.shape[:-1] + (3, 3))


def q_convert_wxyz(q):
    """xyzw -> wxyz."""
    return torch.cat([q[..., 3:4], q[..., 0:3]], dim=-1)


def q_to_euler(q):
    """(w, x, y, z) -> (roll, pitch, yaw) of the Z-Y-X convention, pitch clamped into asin's domain
    (reference: quaternion.py:203-215)."""
    q = q.unsqueeze(0) if q.ndim == 1 else q
    w, x, y, z = torch.unbind(q, -1)
    roll = torch.atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    pitch = torch.asin(torch.clamp(2.0 * (w * y - z * x), -1.0, 1.0))
    yaw = torch.atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return torch.stack([roll, pitch, yaw], dim=-1)
