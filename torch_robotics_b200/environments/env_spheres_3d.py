"""EnvSpheres3D: ten spheres of radius 0.15 in a +-1 m workspace - the obstacle set of BASELINE
config 3 (reference: environments/env_spheres_3d.py:11-49)."""
import torch

from .env_base import EnvBase
from .primitives import MultiSphereField, ObjectField

SPHERE_CENTERS = [
    [-0.3, 0.3, 0.85], [-0.35, -0.25, 0.45], [-0.45, 0.15, 0.0], [0.45, 0.35, 0.0], [0.55, 0.35, 0.55],
    [0.65, -0.4, 0.25], [0.2, -0.35, 0.5], [0.35, 0.0, 0.9], [0.0, -0.3, 0.0], [0.0, 0.45, 0.35],
]
SPHERE_RADIUS = 0.15


class EnvSpheres3D(EnvBase):
    def __init__(self, name="EnvDense2D", tensor_args=None, **kwargs):
        spheres = MultiSphereField(torch.tensor(SPHERE_CENTERS), torch.full((len(SPHERE_CENTERS),), SPHERE_RADIUS),
                                   tensor_args=tensor_args)
        super().__init__(name=name, limits=torch.tensor([[-1, -1, -1], [1, 1, 1]], **tensor_args),
                         obj_fixed_list=[ObjectField([spheres], "spheres")], tensor_args=tensor_args, **kwargs)
