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
