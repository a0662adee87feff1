"""plot_coordinate_frame (reference: visualizers/plot_utils.py:8-32): the three axes of a Frame as red / green / blue arrows."""
import torch

from ..torch_utils.torch_utils import to_numpy


def plot_coordinate_frame(ax, frame, arrow_length=0.1, arrow_alpha=1.0, arrow_linewidth=1.0, tensor_args=None):
    origin = to_numpy(frame.translation).reshape(-1)[:3]
    R = frame.rotation.reshape(-1, 3, 3)[0]
    for col, color in enumerate(("red", "green", "blue")):
        a = to_numpy(R[:, col])
        ax.quiver(origin[0], origin[1], origin[2], a[0], a[1], a[2], length=arrow_length, normalize=True, color=color,
                  alpha=arrow_alpha, linewidth=arrow_linewidth)
