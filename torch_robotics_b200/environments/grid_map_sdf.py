"""GridMapSDF: signed distance and its gradient precomputed on a regular voxel grid, looked up per collision point
(reference: environments/grid_map_sdf.py:9-114; built by EnvBase(precompute_sdf_obj_fixed=True), env_base.py:59-65).

Layout in HBM: one float4 (sdf, d/dx, d/dy, d/dz) per cell, [n0, n1, n2, 4] - the reference keeps two tensors
(sdf_tensor [n0,n1,n2] and grad_sdf_tensor [n0,n1,n2,3]); interleaving them makes a lookup one 16-byte gather.
The grid is filled by the library's own cost-stage kernel run on the grid's sample points (a one-body "point" model,
margin 0: cost = -sdf, d cost/d x = -grad sdf), so no CPU or eager-torch SDF code is involved."""
import numpy as np
import torch

from .. import ops
from .._lib import DeviceCost, DeviceModel
from ..descriptor import VoxelSDF, build_cost_descriptor, point_descriptor


class GridMapSDF:
    def __init__(self, limits, cell_size, obj_list, tensor_args=None, sdf_tensor=None, grad_sdf_tensor=None):
        """sdf_tensor / grad_sdf_tensor: adopt an existing grid (e.g. one computed by the reference) instead of
        precomputing it."""
        self.limits = limits
        self.dim = limits.shape[-1]
        if self.dim != 3:
            raise NotImplementedError("voxel SDFs are implemented for 3-D workspaces (the kernels' collision points are 3-D)")
        self.tensor_args = tensor_args
        self.obj_list = obj_list
        self.map_dim = torch.abs(limits[1] - limits[0])                                   # grid_map_sdf.py:24
        self.cell_size = cell_size
        self.cmap_dim = torch.ceil(self.map_dim / cell_size).type(torch.LongTensor).to(tensor_args["device"])   # :27
        self._n = tuple(int(v) for v in self.cmap_dim.tolist())
        self._point_model = {}
        self._lookup_cost = {}
        if sdf_tensor is not None:
            self.grid = torch.cat((sdf_tensor.unsqueeze(-1), grad_sdf_tensor), dim=-1).to(**tensor_args).contiguous()
            assert tuple(self.grid.shape) == (*self._n, 4)
        else:
            self.grid = None
            self.precompute_sdf()

    # ---- the reference's two views of the grid
    @property
    def sdf_tensor(self):
        return self.grid[..., 0]

    @property
    def grad_sdf_tensor(self):
        return self.grid[..., 1:4]

    @property
    def points_for_sdf(self):
        return torch.stack(torch.meshgrid(*self._axes(), indexing="ij"), dim=-1)

    def _axes(self):
        return [torch.linspace(self.limits[0][a], self.limits[1][a], self._n[a], **self.tensor_args) for a in range(3)]   # :36-44

    def _model(self, device) -> DeviceModel:
        idx = torch.device(device).index or 0
        if idx not in self._point_model:
            self._point_model[idx] = DeviceModel(point_descriptor(), torch.device("cuda", idx))
        return self._point_model[idx]

    def precompute_sdf(self):
        """sdf = min over objects (grid_map_sdf.py:64-72), gradient of the nearest primitive (the reference takes it by
        autograd, :48-60), evaluated slab by slab along the first axis like the reference's loop."""
        device = torch.device(self.tensor_args["device"])
        if device.type != "cuda":
            raise RuntimeError(f"GridMapSDF was requested on '{device}': the SDF grid is computed by the CUDA library only")
        cost = build_cost_descriptor(pt_link=[0], pt_radius=[0.0], cutoff_obj=0.0,
                                     objects=[o.as_sphere_table() for o in self.obj_list], weights=(0.0, 1.0, 0.0, 0.0))
        dc = DeviceCost(self._model(device), cost)
        ax = self._axes()
        n0, n1, n2 = self._n
        self.grid = torch.empty((n0, n1, n2, 4), device=device, dtype=torch.float32)
        yz = torch.stack(torch.meshgrid(ax[1], ax[2], indexing="ij"), dim=-1).reshape(-1, 2)
        slab = torch.empty((n1 * n2, 1, 3), device=device, dtype=torch.float32)
        slab[:, 0, 1:] = yz
        for i in range(n0):
            slab[:, 0, 0] = ax[0][i]
            r = dc.cost_stage(slab, want_grad=True)
            self.grid[i, ..., 0] = -r["total"].view(n1, n2)
            self.grid[i, ..., 1:4] = -r["grad_pos"].view(n1, n2, 3)

    # ---- descriptor side
    def as_sphere_table(self) -> VoxelSDF:
        return VoxelSDF(self._n, self.limits[0].detach().float().cpu().numpy(), self.map_dim.detach().float().cpu().numpy(), self.grid)

    # ---- standalone queries (grid_map_sdf.py:74-112)
    def __call__(self, X, **kwargs):
        return self.get_sdf(X, **kwargs)

    def compute_cost(self, X, **kwargs):
        return self.get_sdf(X, **kwargs)

    def compute_signed_distance(self, X, **kwargs):
        return self.get_sdf(X, **kwargs)

    def get_sdf(self, X, **kwargs):
        """X [..., 3] -> sdf [...] of the cell each point falls into; d sdf / d X is the stored cell gradient (the
        reference's surrogate, :103-112)."""
        idx = X.device.index or 0
        if idx not in self._lookup_cost:
            cost = build_cost_descriptor(pt_link=[0], pt_radius=[0.0], cutoff_obj=0.0, objects=[self.as_sphere_table()],
                                         weights=(0.0, 1.0, 0.0, 0.0))
            self._lookup_cost[idx] = DeviceCost(self._model(X.device), cost)
        return -ops.cost_stage(self._lookup_cost[idx], X.reshape(-1, 1, 3)).reshape(X.shape[:-1])

    def zero_grad(self):
        pass
