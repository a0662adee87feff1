"""precomputed voxel SDF with surrogate gradient (environments/grid_map_sdf.py,
EnvBase(precompute_sdf_obj_fixed=True)).  Golden: tests/golden/planning_panda_voxel_sdf.npz - the reference's own
40^3 grid of EnvSpheres3D plus the Panda objective evaluated through it."""
import numpy as np
import pytest
import torch

from conftest import get_desc, load_golden, rel_err
from oracle import c_oracle as co
from oracle import restatement as orc


def grid4(g):
    return np.ascontiguousarray(np.concatenate([g["sdf_tensor"][..., None], g["grad_sdf_tensor"]], -1), np.float32)


def cost_from_golden(g, data=None, with_rest=True):
    """Objective tables of PlanningTask(EnvSpheres3D(precompute_sdf_obj_fixed=True, obj_extra_list=[extra]), RobotPanda())."""
    from torch_robotics_b200.descriptor import VoxelSDF, build_cost_descriptor
    from torch_robotics_b200.robots.robot_panda import SELF_COLLISION_PAIRS, GRASPED_OBJECT_SELF_LINKS
    lim = g["limits"]
    vox = VoxelSDF(tuple(int(v) for v in g["cmap_dim"]), lim[0], np.abs(lim[1] - lim[0]), grid4(g) if data is None else data)
    kw = {}
    if with_rest:
        idx = get_desc("panda").name_to_idx()
        links = sorted(set(GRASPED_OBJECT_SELF_LINKS) | set(SELF_COLLISION_PAIRS) | {b for v in SELF_COLLISION_PAIRS.values() for b in v})
        pairs = [(idx[a], idx[b]) for a in links for b in SELF_COLLISION_PAIRS.get(a, [])]
        kw = dict(ws_min=lim[0], ws_max=lim[1], self_pairs=pairs, self_margin=np.full(len(pairs), 0.05, np.float32))
    return build_cost_descriptor(pt_link=g["pt_link"], pt_radius=g["pt_radius"], cutoff_obj=float(g["obstacle_cutoff_margin"]),
                                 objects=[vox, (g["extra_centers"], g["extra_radii"], None, None)], **kw)


def test_cell_index_matches_reference():
    from torch_robotics_b200.descriptor import grid_cell_index
    g = load_golden("planning_panda_voxel_sdf")
    cost = cost_from_golden(g)
    pts = g["link_pos"][..., g["pt_link"], :]
    assert np.array_equal(orc.grid_cells(cost, pts), g["cells"])
    assert np.array_equal(grid_cell_index(pts, cost.grid_lo, cost.grid_extent, cost.grid_n).numpy(), g["cells"])
    # outside the workspace: clamped to the border cells
    far = np.array([[-5.0, 0.0, 7.0], [1.0, -1.0, 0.999999]], np.float32)
    assert orc.grid_cells(cost, far).tolist() == [[0, 20, 39], [39, 0, 39]]


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_matches_reference(impl):
    g, d = load_golden("planning_panda_voxel_sdf"), get_desc("panda")
    q = g["q"].reshape(-1, 7)
    for which, cost in (("obj", cost_from_golden(g, with_rest=False)), ("collision", cost_from_golden(g))):
        o = orc.objective(d, cost, q, dtype=np.float32) if impl == "numpy" else co.objective(co.Tree(d), co.Cost(cost), q)
        assert rel_err(o["total"], g[f"cost_{which}"].reshape(-1)) < 1e-5
        assert rel_err(o["grad_q"], g[f"grad_{which}"].reshape(-1, 7)) < 1e-4
        # the oracle reads the reference's own grid, so only points within an ulp of a cell face could differ
        want = g["occ_obj" if which == "obj" else "occ_all"].reshape(-1)
        got = ((o["occ"] >> 1) & 1) if which == "obj" else (o["occ"] != 0)
        assert np.mean(got != want) < 0.01
        cells = g["cells"].reshape(q.shape[0], -1, 3).astype(np.int64)
        lin = (cells[..., 0] * 40 + cells[..., 1]) * 40 + cells[..., 2]
        assert np.mean(o["argmin"][:, 0, :] != lin) < 0.01


def test_descriptor_rejects_two_grids():
    from torch_robotics_b200.descriptor import VoxelSDF, build_cost_descriptor
    v = VoxelSDF((2, 2, 2), np.zeros(3), np.ones(3), np.zeros((2, 2, 2, 4), np.float32))
    with pytest.raises(ValueError):
        build_cost_descriptor(pt_link=[0], pt_radius=[0.1], objects=[v, v])


# ---------------------------------------------------------------------------------------------- GPU
TA = dict(device="cuda:0", dtype=torch.float32)


def make_env(g, **kw):
    from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D
    from torch_robotics_b200.environments.primitives import MultiSphereField, ObjectField
    extra = ObjectField([MultiSphereField(torch.tensor(g["extra_centers"]), torch.tensor(g["extra_radii"]), tensor_args=TA)], "extra")
    return EnvSpheres3D(precompute_sdf_obj_fixed=True, sdf_cell_size=float(g["cell_size"]), obj_extra_list=[extra], tensor_args=TA, **kw)


@pytest.mark.gpu
def test_gpu_precomputed_grid_matches_reference():
    """The grid filled by the cost-stage kernel equals the reference's (torch SDF + autograd gradient)."""
    g = load_golden("planning_panda_voxel_sdf")
    grid = make_env(g).grid_map_sdf_obj_fixed
    assert tuple(grid.cmap_dim.tolist()) == tuple(g["cmap_dim"])
    assert rel_err(grid.sdf_tensor.cpu().numpy(), g["sdf_tensor"]) < 1e-5
    # gradient: unit vectors; cells on a bisector between two spheres may pick the other one
    dg = np.abs(grid.grad_sdf_tensor.cpu().numpy() - g["grad_sdf_tensor"]).max(-1)
    assert np.mean(dg > 1e-5) < 1e-3
    assert rel_err(grid.points_for_sdf[3, 7, 11].cpu().numpy(), np.array([-1 + 2 * 3 / 39, -1 + 2 * 7 / 39, -1 + 2 * 11 / 39])) < 1e-6


@pytest.mark.gpu
def test_gpu_planning_through_grid_matches_reference():
    """PlanningTask on an environment with a precomputed grid, adopting the reference's grid so lookups are comparable."""
    from torch_robotics_b200.environments.grid_map_sdf import GridMapSDF
    from torch_robotics_b200.robots.robot_panda import RobotPanda
    from torch_robotics_b200.tasks.tasks import PlanningTask
    g = load_golden("planning_panda_voxel_sdf")
    env = make_env(g)
    env.grid_map_sdf_obj_fixed = GridMapSDF(env.limits, float(g["cell_size"]), env.obj_fixed_list, tensor_args=TA,
                                            sdf_tensor=torch.tensor(g["sdf_tensor"]), grad_sdf_tensor=torch.tensor(g["grad_sdf_tensor"]))
    task = PlanningTask(env=env, robot=RobotPanda(tensor_args=TA), tensor_args=TA)
    q = torch.tensor(g["q"], device="cuda:0").requires_grad_(True)
    pos = task.robot.fk_map_collision(q)
    c_obj = task.df_collision_objects.compute_cost(q, pos, field_type="sdf")
    assert rel_err(c_obj.detach().cpu().numpy(), g["cost_obj"]) < 1e-5
    c_obj.sum().backward()
    assert rel_err(q.grad.cpu().numpy(), g["grad_obj"]) < 1e-4
    q.grad = None
    c = task.compute_collision_cost(q)
    assert rel_err(c.detach().cpu().numpy(), g["cost_collision"]) < 1e-5
    c.sum().backward()
    assert rel_err(q.grad.cpu().numpy(), g["grad_collision"]) < 1e-4
    with torch.no_grad():
        # stage-wise: on the reference's own link positions the masks are bit-exact
        pos_ref = torch.tensor(g["link_pos"], device="cuda:0")
        for key, kw in (("occ_obj", {}), ("occ_obj_zero", {"margin": 0.0})):
            occ = task.df_collision_objects.compute_cost(q, pos_ref, field_type="occupancy", **kw)
            assert np.array_equal(occ.cpu().numpy(), g[key])
        assert np.mean(task.compute_collision(q).cpu().numpy() != g["occ_all"]) < 0.01
    # standalone queries, inside and outside the workspace
    X = torch.tensor(g["query_x"], device="cuda:0").requires_grad_(True)
    sdf = env.grid_map_sdf_obj_fixed(X)
    assert np.array_equal(sdf.detach().cpu().numpy(), g["query_sdf"])
    sdf.sum().backward()
    assert np.array_equal(X.grad.cpu().numpy(), g["query_grad"])


@pytest.mark.gpu
def test_gpu_bit_exact_vs_oracle_at_scale():
    """Fused objective through the library's own grid (cell 0.02 -> 100^3) vs the C oracle reading the same grid:
    costs 1e-5, gradients 1e-4, occupancy bits and looked-up cell indices bit-exact on 200k configurations."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D
    g, d = load_golden("planning_panda_voxel_sdf"), get_desc("panda")
    env = EnvSpheres3D(precompute_sdf_obj_fixed=True, sdf_cell_size=0.02, tensor_args=TA)
    vox = env.grid_map_sdf_obj_fixed.as_sphere_table()
    assert vox.n == (100, 100, 100)
    cost = cost_from_golden(g)
    cost.grid_n, cost.grid_data = np.asarray(vox.n, np.int32), vox.data
    rng = np.random.default_rng(5)
    lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float32)
    hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float32)
    q = (lo + rng.random((200_000, 7), dtype=np.float32) * (hi - lo)).astype(np.float32)
    m = DeviceModel(d, "cuda:0")
    assert m.kernel_family == "panda"
    out = DeviceCost(m, cost).objective(torch.as_tensor(q, device="cuda:0"), want_grad=True, want_occupancy=True, want_argmin=True)
    o = co.objective(co.Tree(d), co.Cost(cost), q)
    # link positions agree to 1e-7 between the two FK implementations, so a point within an ulp of a cell face may
    # land in the neighbouring cell: rare rows; the lookup itself is pinned bit-exactly below on shared positions
    assert np.mean(out["occupancy"].cpu().numpy() != o["occ"]) < 1e-4
    assert np.mean(out["argmin"].cpu().numpy() != o["argmin"]) < 1e-4
    same = (out["argmin"].cpu().numpy() == o["argmin"]).all(axis=(1, 2))     # the field is piecewise constant per cell
    assert rel_err(out["total"].cpu().numpy()[same], o["total"][same]) < 1e-5
    gd = np.abs(out["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1)[same] / np.abs(o["grad_q"]).max()
    assert (gd >= 1e-4).sum() <= 30 and gd[gd < 1e-4].max() < 1e-4
    # stage-wise: the same device positions through the cost-stage kernel and the numpy restatement
    n = 20_000
    pos = m.fk_forward(torch.as_tensor(q[:n], device="cuda:0"))[..., :3, 3].contiguous()
    st = DeviceCost(m, cost).cost_stage(pos, want_grad=False, want_occupancy=True, want_argmin=True)
    ref = orc.object_term(cost, pos.cpu().numpy()[:, cost.pt_link, :])
    assert np.array_equal(st["argmin"].cpu().numpy(), ref["argmin"])
    assert np.array_equal((st["occupancy"].cpu().numpy() >> 1) & 1, ref["occ"].astype(np.uint8))
