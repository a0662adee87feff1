"""Grasped objects (reference: environments/objects.py:10-89): an ObjectField posed relative to the end-effector
link whose collision is checked through a fixed set of surface points carried by the kinematic tree."""
import abc

import torch

from .primitives import MultiBoxField, ObjectField


class GraspedObject(ObjectField):
    def __init__(self, primitive_fields, **kwargs):
        assert len(primitive_fields) == 1, "a grasped object has exactly one primitive (objects.py:16)"
        super().__init__(primitive_fields, **kwargs)
        if not isinstance(self.fields[0], MultiBoxField):
            raise NotImplementedError("only box-shaped grasped objects exist in the reference (objects.py:24-31)")

    @abc.abstractmethod
    def get_base_points_for_collision(self):
        raise NotImplementedError


class GraspedObjectPandaBox(GraspedObject):
    """5 x 5 x 15 cm box held 11 cm in front of panda_hand, long side across the fingers (objects.py:37-57)."""

    def __init__(self, tensor_args=None, **kwargs):
        box = MultiBoxField(torch.zeros(1, 3, **tensor_args), torch.tensor([[0.05, 0.05, 0.15]], **tensor_args),
                            tensor_args=tensor_args)
        super().__init__([box], name="GraspedObjectPandaBox",
                         pos=torch.tensor([0., 0., 0.11], **tensor_args),
                         ori=torch.tensor([0, 0.7071081, 0, 0.7071055], **tensor_args),
                         reference_frame="panda_hand", **kwargs)
        self.base_points_for_collision = self.get_base_points_for_collision()
        self.n_base_points_for_collision = len(self.base_points_for_collision)

    def get_base_points_for_collision(self):
        """The 8 corners (bottom ring, then top ring, each starting at +x+y and going -y, -x, +y) followed by the
        6 face centres in the order +x, -y, -x, +y, +z, -z (objects.py:59-89)."""
        hx, hy, hz = (self.fields[0].sizes[0] / 2).tolist()
        ring = [(hx, hy), (hx, -hy), (-hx, -hy), (-hx, hy)]
        corners = [[x, y, z] for z in (-hz, hz) for x, y in ring]
        faces = [[hx, 0, 0], [0, -hy, 0], [-hx, 0, 0], [0, hy, 0], [0, 0, hz], [0, 0, -hz]]
        return torch.tensor(corners + faces, **self.tensor_args)
