"""Waypoint interpolation used by the trajectory collision check (reference: trajectory/utils.py:37-50)."""
import torch


def interpolate_traj_via_points(trajs, num_interpolation=10):
    """Linear interpolation between consecutive waypoints.  For [..., H, D] input returns the
    (H-1) * num_interpolation interior points only (the waypoints themselves are not repeated),
    ordered segment by segment; within a segment the blend weight of the segment's first waypoint runs
    over linspace(0, 1, num_interpolation + 2)[1:-1], as in the reference."""
    if num_interpolation <= 0:
        return trajs
    assert trajs.ndim > 1
    D = trajs.shape[-1]
    alpha = torch.linspace(0, 1, num_interpolation + 2).type_as(trajs)[1:num_interpolation + 1].view(-1, 1)
    a, b = trajs[..., :-1, None, :], trajs[..., 1:, None, :]
    out = a * alpha + b * (1 - alpha)
    return out.reshape(trajs.shape[:-2] + (-1, D))


def finite_difference_vector(x, dt=1., method="forward"):
    """Time derivative of a trajectory x [..., H, D] by finite differences along the horizon; border rows that have no
    neighbour on the needed side stay zero (reference: trajectory/utils.py:53-64)."""
    out = torch.zeros_like(x)
    if method == "forward":
        out[..., :-1, :] = (x[..., 1:, :] - x[..., :-1, :]) / dt
    elif method == "backward":
        out[..., 1:, :] = (x[..., 1:, :] - x[..., :-1, :]) / dt
    elif method == "central":
        out[..., 1:-1, :] = (x[..., 2:, :] - x[..., :-2, :]) / (2 * dt)
    else:
        raise NotImplementedError(method)
    return out
