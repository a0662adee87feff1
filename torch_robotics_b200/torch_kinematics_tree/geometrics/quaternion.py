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


def rotation_matrix_to_q(rot_mat):
    """Rotation matrices [..., 3, 3] -> unit quaternions (w, x, y, z).  Of the four algebraically equal candidates the one
    built around the largest of sqrt(1 +- m00 +- m11 +- m22) is returned (its own component is then positive and the
    division is well conditioned; first candidate on ties) - what the reference computes at geometrics/quaternion.py:135-166."""
    lead = rot_mat.shape[:-2]
    m = rot_mat.reshape(lead + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    t = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1)
    q_abs = torch.sqrt(torch.clamp(t, min=0.0))
    pick = q_abs.argmax(dim=-1, keepdim=True)                       # [..., 1]
    sq = q_abs * q_abs
    cand = torch.stack([
        torch.stack([sq[..., 0], m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, sq[..., 1], m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, sq[..., 2], m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, sq[..., 3]], dim=-1)], dim=-2)    # [..., 4 candidates, 4]
    row = cand.gather(-2, pick[..., None].expand(lead + (1, 4)))[..., 0, :]
    den = 2.0 * torch.clamp(q_abs.gather(-1, pick), min=0.1)
    return row / den
