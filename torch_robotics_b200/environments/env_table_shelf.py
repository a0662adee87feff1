"""EnvTableShelf: a table in front of the robot and a ten-board shelf next to it, all rounded boxes (reference:
environments/env_table_shelf.py:14-102).  Geometry is kept as dimension tables; the resulting fp32 centres / sizes /
poses are bit-identical to the reference's (tests/golden/planning_panda_table_shelf.npz)."""
import itertools

import numpy as np
import torch

from .env_base import EnvBase
from .primitives import MultiBoxField, ObjectField

TABLE_SIZE = (0.56, 0.90, 0.80)                      # x, y, z extents in the table frame
SHELF = dict(width=0.80, height=2.05, depth=0.28, panel=0.02, board=0.015,
             first_board_z=0.82, board_gaps=(0.23, 0.255, 0.225, 0.225))
ROBOT_TO_TABLE = 0.10
TABLE_TO_SHELF = 0.15


def table_object(tensor_args=None):
    box = MultiBoxField(np.array([(0., 0., 0.)]), np.array([TABLE_SIZE]), tensor_args=tensor_args)
    return ObjectField([box], "table")


def shelf_object(tensor_args=None):
    """Boxes in the shelf's own frame (origin at the outer bottom corner of the left panel): two side panels, the back
    panel, bottom and top boards, then five boards starting at first_board_z with the listed gaps."""
    s = SHELF
    w, h, d, p, t = s["width"], s["height"], s["depth"], s["panel"], s["board"]
    inner = w - 2 * p
    mid_x = p + inner / 2
    board = (inner, d, t)
    boxes = [((p / 2, d / 2, h / 2), (p, d, h)),
             ((p + inner + p / 2, d / 2, h / 2), (p, d, h)),
             ((mid_x, d + p / 2, h / 2), (inner, p, h)),
             ((mid_x, d / 2, t / 2), board),
             ((mid_x, d / 2, h - t / 2), board)]
    for z in itertools.accumulate(s["board_gaps"], initial=s["first_board_z"] + t / 2):
        boxes.append(((mid_x, d / 2, z), board))
    centers = np.array([c for c, _ in boxes])
    sizes = np.array([sz for _, sz in boxes])
    return ObjectField([MultiBoxField(centers, sizes, tensor_args=tensor_args)], "shelf")


class EnvTableShelf(EnvBase):
    def __init__(self, tensor_args=None, **kwargs):
        table = table_object(tensor_args)
        tx, ty, tz = (v.item() for v in table.fields[0].sizes[0])
        quarter_turn = np.deg2rad(90)                                   # the table's long side faces the robot
        table.set_position_orientation(pos=(ROBOT_TO_TABLE + ty / 2, 0, -tz / 2),
                                       ori=[np.cos(quarter_turn / 2), 0, 0, np.sin(quarter_turn / 2)])
        shelf = shelf_object(tensor_args)
        shelf.set_position_orientation(pos=(ROBOT_TO_TABLE, TABLE_TO_SHELF + tx / 2, -tz))
        super().__init__(name=self.__class__.__name__, limits=torch.tensor([[-1, -1, -1], [1.5, 1., 1.5]], **tensor_args),
                         obj_fixed_list=[table, shelf], tensor_args=tensor_args, **kwargs)
