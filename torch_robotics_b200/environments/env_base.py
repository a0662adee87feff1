"""EnvBase: workspace limits + obstacle object lists (reference: environments/env_base.py:17-88).
Only what feeds the distance fields is kept - including the optional precomputed voxel SDF of the fixed objects
(env_base.py:57-65); occupancy maps and rendering are outside the hot path."""
import itertools
from abc import ABC

from ..torch_utils.torch_utils import to_numpy
from .primitives import ObjectField


class EnvBase(ABC):
    def __init__(self, name="NameEnvBase", limits=None, obj_fixed_list=None, obj_extra_list=None,
                 precompute_sdf_obj_fixed=False, sdf_cell_size=0.005, tensor_args=None, **kwargs):
        assert limits is not None
        self.tensor_args = tensor_args
        self.name = name
        self.limits = limits
        self.limits_np = to_numpy(limits)
        self.dim = len(limits[0])
        for lst in (obj_fixed_list, obj_extra_list):
            for obj in lst or []:
                assert isinstance(obj, ObjectField), "Objects must be instances of ObjectField class"
        self.obj_fixed_list = obj_fixed_list
        self.obj_extra_list = obj_extra_list
        self.obj_all_list = set(itertools.chain(obj_fixed_list or [], obj_extra_list or []))
        self.grid_map_sdf_obj_fixed = None
        if precompute_sdf_obj_fixed:
            from .grid_map_sdf import GridMapSDF
            self.grid_map_sdf_obj_fixed = GridMapSDF(self.limits, sdf_cell_size, self.obj_fixed_list, tensor_args=tensor_args)
        self.occupancy_map = None

    def get_obj_list(self):
        return self.obj_all_list

    def get_df_obj_list(self, return_extra_objects_only=False):
        objs = []
        if not return_extra_objects_only:
            if self.grid_map_sdf_obj_fixed is not None:      # the grid stands in for all fixed objects (env_base.py:80-84)
                objs.append(self.grid_map_sdf_obj_fixed)
            else:
                objs.extend(self.obj_fixed_list or [])
        if self.obj_extra_list is not None:
            objs.extend(self.obj_extra_list)
        return objs

    def zero_grad(self):
        for obj in self.obj_all_list:
            obj.zero_grad()
