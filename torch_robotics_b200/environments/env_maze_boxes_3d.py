"""EnvMazeBoxes3D: fourteen 1.95 m tall rounded-box pillars forming four walls with gaps (reference:
environments/env_maze_boxes_3d.py:11-65).  The footprints are kept as (x_min, y_min, x_max, y_max) rectangles."""
import numpy as np
import torch

from .env_base import EnvBase
from .primitives import MultiBoxField, ObjectField

PILLAR_HEIGHT = 1.95
_T = 0.2      # wall thickness


def _wall(x_min, x_max, spans):
    return [(x_min, y0, x_max, y1) for y0, y1 in spans]


FOOTPRINTS = (
    _wall(-0.75, -0.75 + _T, [(-1.0, -0.3), (0.3, 1.0), (-0.15, 0.15)])
    + _wall(-_T / 2 - 0.2, _T / 2 - 0.2, [(-1, -1 + 0.05), (-1 + 0.2, 1 - 0.2), (1 - 0.05, 1)])
    + _wall(-_T / 2 + 0.2, _T / 2 + 0.2, [(-1.0, -0.3), (0.3, 1.0), (-0.15, 0.15)])
    + _wall(0.75 - _T, 0.75, [(-1, -0.6), (-0.5, -0.2), (-0.1, 0.1), (0.2, 0.5), (0.6, 1)])
)


def pillar_objects(tensor_args=None):
    centers = np.array([(x0 + abs(x1 - x0) / 2, y0 + abs(y1 - y0) / 2, 0) for x0, y0, x1, y1 in FOOTPRINTS])
    sizes = np.array([(abs(x1 - x0), abs(y1 - y0), PILLAR_HEIGHT) for x0, y0, x1, y1 in FOOTPRINTS])
    return [ObjectField([MultiBoxField(centers, sizes, tensor_args=tensor_args)], "boxes")]


class EnvMazeBoxes3D(EnvBase):
    def __init__(self, tensor_args=None, **kwargs):
        super().__init__(name=self.__class__.__name__, limits=torch.tensor([[-1, -1, -1], [1, 1, 1]], **tensor_args),
                         obj_fixed_list=pillar_objects(tensor_args), tensor_args=tensor_args, **kwargs)
