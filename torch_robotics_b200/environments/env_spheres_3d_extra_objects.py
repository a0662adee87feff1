"""EnvSpheres3DExtraObjects: EnvSpheres3D plus five more spheres held in a separate 'extra' object, i.e. obstacles that
were not there when a model was trained (reference: environments/env_spheres_3d_extra_objects.py:12-58)."""
import numpy as np

from .env_spheres_3d import SPHERE_RADIUS, EnvSpheres3D
from .primitives import MultiSphereField, ObjectField

EXTRA_SPHERE_CENTERS = [[0.25, 0., 0.], [0., 0.5, 0.5], [0., -0.5, 0.], [-0.25, -0.5, 0.5], [-0.25, 0., 0.75]]


class EnvSpheres3DExtraObjects(EnvSpheres3D):
    def __init__(self, tensor_args=None, **kwargs):
        extra = MultiSphereField(np.array(EXTRA_SPHERE_CENTERS), np.full(len(EXTRA_SPHERE_CENTERS), SPHERE_RADIUS),
                                 tensor_args=tensor_args)
        super().__init__(name=self.__class__.__name__, obj_extra_list=[ObjectField([extra], "extra-objects")],
                         tensor_args=tensor_args, **kwargs)
