"""Free functions callers of the hot path use (reference: geometrics/utils.py:11-17,130-154,209-224,
321-344).  SE3_distance here is the eager torch form for API parity on arbitrary tensors; inside the
planning objective and the IK loss the same quantity is computed by the fused CUDA kernel."""
import torch


def multiply_transform(w_rot_l, w_trans_l, l_rot_c, l_trans_c):
    w_rot_c = w_rot_l @ l_rot_c
    w_trans_c = (w_rot_l @ l_trans_c.float().unsqueeze(-1)).squeeze(-1) + w_trans_l
    return w_rot_c, w_trans_c


def so3_rotation_angle(R, cos_angle=False, eps=1e-4):
    if R.shape[-2:] != (3, 3):
        raise ValueError("Input has to be a batch of 3x3 Tensors.")
    phi_cos = (R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2] - 1.0) * 0.5
    if cos_angle:
        return phi_cos
    bound = 1.0 - eps
    return torch.acos(phi_cos.clamp(-bound, bound)) if eps > 0 else torch.acos(phi_cos)


def so3_relative_angle(R1, R2, cos_angle=False, eps=1e-4):
    return so3_rotation_angle(R1 @ R2.transpose(-2, -1), cos_angle=cos_angle, eps=eps)


def SE3_distance(H_batch, H_target, w_pos=1.0, w_rot=1.0, broadcast_target=False, **unused):
    """w_rot * (1 - cos(relative angle)) + w_pos * |p - p_target|."""
    if H_batch.ndim == 2:
        H_batch = H_batch.unsqueeze(0)
    if H_target.ndim == 2:
        H_target = H_target.unsqueeze(0)
    if broadcast_target:
        H_target = H_target.unsqueeze(1)
    D = 0.0
    if w_rot > 0.0:
        D = D + w_rot * (1 - so3_relative_angle(H_batch[..., :3, :3], H_target[..., :3, :3], cos_angle=True))
    if w_pos > 0.0:
        D = D + w_pos * torch.linalg.norm(H_batch[..., :3, 3] - H_target[..., :3, 3], dim=-1)
    return D


def link_pos_from_link_tensor(link_tensor):
    if link_tensor.shape[-1] == 3:
        return link_tensor[..., :2, 2]
    if link_tensor.shape[-1] == 4:
        return link_tensor[..., :3, 3]
    raise ValueError


def link_rot_from_link_tensor(link_tensor):
    if link_tensor.shape[-1] == 3:
        return link_tensor[..., :2, :2]
    if link_tensor.shape[-1] == 4:
        return link_tensor[..., :3, :3]
    raise ValueError


def link_quat_from_link_tensor(link_tensor):
    """wxyz quaternions of the links' rotations (reference: geometrics/utils.py:341-344)."""
    from .quaternion import rotation_matrix_to_q
    return rotation_matrix_to_q(link_rot_from_link_tensor(link_tensor))
