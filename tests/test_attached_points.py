"""collision points that are not link origins - the grasped box of
environments/objects.py and the per-link multi-sphere model of data/configs/panda - realised as fixed bodies attached
to the kinematic tree.  Goldens: tests/golden/points_*.npz (make_golden.py:golden_attached_points)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import c_oracle as co
from oracle import restatement as orc

TA_CPU = dict(device="cpu", dtype=torch.float32)


def grasped_robot(tensor_args):
    from torch_robotics_b200.environments.objects import GraspedObjectPandaBox
    from torch_robotics_b200.robots.robot_panda import RobotPanda
    return RobotPanda(grasped_object=GraspedObjectPandaBox(tensor_args=tensor_args), tensor_args=tensor_args)


def sphere_robot(tensor_args):
    from torch_robotics_b200.robots.robot_panda import RobotPanda
    return RobotPanda(sphere_config=True, tensor_args=tensor_args)


def cost_of(robot, g, with_self):
    from torch_robotics_b200.descriptor import build_cost_descriptor
    kw = {}
    if with_self:
        f = robot.df_collision_self
        rows = list(f.link_idxs_for_collision_checking)
        kw = dict(self_pairs=[(rows[a], rows[b]) for a, b in f.idxs_links_distance_matrix],
                  self_margin=f.cutoff_margin.numpy())
    return build_cost_descriptor(
        pt_link=robot.point_idxs_for_object_collision_checking,
        pt_radius=robot.link_margins_for_object_collision_checking_tensor.numpy(), cutoff_obj=float(g["cutoff"]),
        objects=[(g["centers"], g["radii"], None, None)], ws_min=g["ws_min"], ws_max=g["ws_max"], **kw)


def test_grasped_object_model_matches_reference_setup():
    g = load_golden("points_panda_grasped_box")
    r = grasped_robot(TA_CPU)
    tree = r.diff_panda
    assert tree.get_link_names() == list(g["link_names"])                 # 11 links + grasped_object
    assert np.array_equal(r.grasped_object.base_points_for_collision.numpy(), g["base_points"])
    assert np.array_equal(r.grasped_object.pos.numpy(), g["grasp_pos"])
    assert np.array_equal(r.grasped_object.ori.numpy(), g["grasp_ori"])
    assert r.point_idxs_for_object_collision_checking == list(g["pt_cols"])
    assert np.array_equal(r.link_margins_for_object_collision_checking_tensor.numpy(), g["pt_radius"])
    f = r.df_collision_self
    assert list(f.link_idxs_for_collision_checking) == list(g["self_rows"])
    assert np.array_equal(np.array(f.idxs_links_distance_matrix, np.int32), g["self_pairs_local"])
    assert np.array_equal(f.cutoff_margin.numpy(), g["self_margin"])
    assert len(tree.get_all_body_names()) == g["link_pos"].shape[-2] == 26


def test_sphere_model_matches_reference_yaml():
    g = load_golden("points_panda_sphere_model")
    r = sphere_robot(TA_CPU)
    d = r.diff_panda._desc
    ids = r.point_idxs_for_object_collision_checking
    assert ids == list(g["pt_cols"]) and len(ids) == 45
    assert [d.link_names[d.parent[i]] for i in ids] == list(g["sphere_links"])
    assert np.array_equal(d.trans[ids], g["sphere_offsets"])
    assert np.array_equal(r.link_margins_for_object_collision_checking_tensor.numpy(), g["pt_radius"])
    assert r.diff_panda.get_link_names() == d.link_names[:11]             # points are hidden from the link list


@pytest.mark.parametrize("case", ["grasped_box", "sphere_model"])
def test_oracle_matches_reference(case):
    g = load_golden(f"points_panda_{case}")
    r = grasped_robot(TA_CPU) if case == "grasped_box" else sphere_robot(TA_CPU)
    d = r.diff_panda._desc
    cost = cost_of(r, g, case == "grasped_box")
    q = g["q"].reshape(-1, 7)
    for o in (orc.objective(d, cost, q, dtype=np.float32), co.objective(co.Tree(d), co.Cost(cost), q)):
        if "pw" in o:
            assert rel_err(o["pw"], g["link_pos"].reshape(q.shape[0], -1, 3)) < 1e-5
        assert rel_err(o["terms"][:, 1], g["cost_obj"].reshape(-1)) < 1e-5
        assert rel_err(o["terms"][:, 2], g["cost_ws"].reshape(-1)) < 1e-5
        assert rel_err(o["total"], g["cost_total"].reshape(-1)) < 1e-5
        assert rel_err(o["grad_q"], g["grad_total"].reshape(-1, 7)) < 1e-4
        if case == "grasped_box":
            assert rel_err(o["terms"][:, 0], g["cost_self"].reshape(-1)) < 1e-5
            assert np.array_equal(o["occ"] & 1, g["occ_self"].reshape(-1).astype(np.uint8))
        # occupancy decisions are pinned stage-wise elsewhere; against the torch reference allow boundary rows
        assert np.mean(((o["occ"] >> 1) & 1) != g["occ_obj"].reshape(-1)) < 0.01
        assert np.mean(((o["occ"] >> 2) & 1) != g["occ_ws"].reshape(-1)) < 0.01


def test_attach_errors():
    from torch_robotics_b200.descriptor import attach_points
    from conftest import get_desc
    d = get_desc("panda")
    with pytest.raises(ValueError):
        attach_points(d, [("no_such_link", (0, 0, 0))])
    with pytest.raises(ValueError):
        attach_points(d, [("panda_hand", (0, 0, 0))] * 60)               # 11 + 60 > TRK_MAX_LINKS
    d2, ids = attach_points(d, [("panda_hand", (0.1, 0, 0)), (3, (0, 0.2, 0))])
    assert ids == [11, 12] and d2.n_links == 13 and d.n_links == 11       # the input descriptor is untouched
    assert list(d2.parent[ids]) == [9, 3] and d2.n_dofs == 7


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["grasped_box", "sphere_model"])
def test_gpu_planning_task_matches_reference(case):
    """Through the reference-facing API (PlanningTask / fields / autograd) on the GPU."""
    from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D
    from torch_robotics_b200.tasks.tasks import PlanningTask
    TA = dict(device="cuda:0", dtype=torch.float32)
    g = load_golden(f"points_panda_{case}")
    r = grasped_robot(TA) if case == "grasped_box" else sphere_robot(TA)
    assert r.diff_panda.device_model("cuda:0").kernel_family == "generic"   # not a compiled-in tree any more
    task = PlanningTask(env=EnvSpheres3D(tensor_args=TA), robot=r, obstacle_cutoff_margin=float(g["cutoff"]), tensor_args=TA)
    q = torch.as_tensor(g["q"], device="cuda:0").requires_grad_(True)
    pos = r.fk_map_collision(q)
    assert rel_err(pos.detach().cpu().numpy(), g["link_pos"]) < 1e-5
    c_obj = task.df_collision_objects.compute_cost(q, pos, field_type="sdf")
    c_ws = task.df_collision_ws_boundaries.compute_cost(q, pos, field_type="sdf")
    assert rel_err(c_obj.detach().cpu().numpy(), g["cost_obj"]) < 1e-5
    assert rel_err(c_ws.detach().cpu().numpy(), g["cost_ws"]) < 1e-5
    if case == "grasped_box":
        c_self = task.df_collision_self.compute_cost(q, pos, field_type="sdf")
        assert rel_err(c_self.detach().cpu().numpy(), g["cost_self"]) < 1e-5
        total = task.compute_collision_cost(q)                                # fused kernel: self + obj + ws
        assert rel_err(total.detach().cpu().numpy(), g["cost_total"]) < 1e-5
        total.sum().backward()
        assert rel_err(q.grad.cpu().numpy(), g["grad_total"]) < 1e-4
        occ = task.df_collision_self.compute_cost(q, pos, field_type="occupancy")
        assert np.array_equal(occ.cpu().numpy(), g["occ_self"])
    else:                                                                     # no self model in this golden: staged path
        (c_obj + c_ws).sum().backward()
        assert rel_err(q.grad.cpu().numpy(), g["grad_obj"] + g["grad_ws"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["grasped_box", "sphere_model"])
def test_gpu_bit_exact_masks_vs_oracle(case):
    """Decisions (occupancy bits) are bit-identical to the C oracle on 200k random configurations."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    g = load_golden(f"points_panda_{case}")
    r = grasped_robot(TA_CPU) if case == "grasped_box" else sphere_robot(TA_CPU)
    d = r.diff_panda._desc
    cost = cost_of(r, g, case == "grasped_box")
    rng = np.random.default_rng(77)
    q = (r.q_min.numpy() + rng.random((200_000, 7), dtype=np.float32) * (r.q_max.numpy() - r.q_min.numpy())).astype(np.float32)
    dc = DeviceCost(DeviceModel(d, "cuda:0"), cost)
    out = dc.objective(torch.as_tensor(q, device="cuda:0"), want_grad=True, want_terms=True, want_occupancy=True)
    o = co.objective(co.Tree(d), co.Cost(cost), q)
    assert np.array_equal(out["occupancy"].cpu().numpy(), o["occ"])
    assert rel_err(out["total"].cpu().numpy(), o["total"]) < 1e-5
    gd = np.abs(out["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    assert (gd >= 1e-4).sum() <= 30 and gd[gd < 1e-4].max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["grasped_box", "sphere_model"])
@pytest.mark.parametrize("scene", ["table_shelf", "spheres3d_extra_voxel"])
def test_gpu_attached_models_on_general_scenes_vs_oracle(case, scene):
    """Attached-body Panda models on scenes beyond the single sphere object - posed rounded boxes; a voxel SDF plus an extra
    sphere object - run on k_objective_ext<..., SCENE = 1> (general_scene_eval).  Terms, gradients, occupancy and per-object
    arg-min against the C oracle; and against the runtime-tree kernels of the same library (TRK_B200_NO_HOT=1 in a fresh
    process is not needed: the oracle is the reference of both)."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    from torch_robotics_b200.environments.env_spheres_3d_extra_objects import EnvSpheres3DExtraObjects
    from torch_robotics_b200.environments.env_table_shelf import EnvTableShelf
    from torch_robotics_b200.tasks.tasks import PlanningTask
    TA = dict(device="cuda:0", dtype=torch.float32)
    r = grasped_robot(TA) if case == "grasped_box" else sphere_robot(TA)
    env = EnvTableShelf(tensor_args=TA) if scene == "table_shelf" else \
        EnvSpheres3DExtraObjects(precompute_sdf_obj_fixed=True, sdf_cell_size=0.04, tensor_args=TA)
    task = PlanningTask(env=env, robot=r, tensor_args=TA)
    cost = task.cost_descriptor()
    d = r.diff_panda._desc
    rng = np.random.default_rng(19)
    lo, hi = r.q_min.cpu().numpy(), r.q_max.cpu().numpy()
    q = (lo + rng.random((20_011, 7), dtype=np.float32) * (hi - lo)).astype(np.float32)
    dc = DeviceCost(DeviceModel(d, "cuda:0"), cost)
    out = dc.objective(torch.as_tensor(q, device="cuda:0"), want_grad=True, want_terms=True, want_occupancy=True, want_argmin=True)
    if getattr(cost, "grid_obj", -1) >= 0:      # the oracle reads the same grid
        cost.grid_data = cost.grid_data.detach().cpu().numpy() if torch.is_tensor(cost.grid_data) else cost.grid_data
    o = co.objective(co.Tree(d), co.Cost(cost), q)
    assert rel_err(out["terms"].cpu().numpy(), o["terms"]) < 1e-5
    same = np.all((out["argmin"].cpu().numpy() == o["argmin"]).reshape(len(q), -1), axis=1)
    assert np.mean(same) > 0.998
    gd = np.abs(out["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    assert np.all(gd[same] < 1e-4)
    assert np.mean(out["occupancy"].cpu().numpy() == o["occ"]) > 0.9995
    occ0 = dc.objective(torch.as_tensor(q, device="cuda:0"), want_grad=False, want_terms=False, want_occupancy=True, margin_override=0.0)
    o0 = co.objective(co.Tree(d), co.Cost(cost), q, margin_override=0.0)
    assert np.mean(occ0["occupancy"].cpu().numpy() == o0["occ"]) > 0.9995


# ---------------------------------------------------------------------------------------------- Tiago sphere model
def tiago_with_spheres():
    """tiago_dual_holo (19 links, two arms) + the 14 spheres of data/configs/tiago, attached as points."""
    import os
    from conftest import get_desc
    from torch_robotics_b200.descriptor import attach_points, build_cost_descriptor, load_sphere_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spheres = load_sphere_config(os.path.join(root, "torch_robotics_b200", "data", "configs", "tiago", "tiago_sphere_config.json"))
    d, ids = attach_points(get_desc("tiago_dual_holo"), [(s[0], s[1]) for s in spheres])
    return d, ids, spheres, build_cost_descriptor


def tiago_cost(g, ids, spheres, build):
    return build(pt_link=ids, pt_radius=np.array([s[2] for s in spheres], np.float32), cutoff_obj=float(g["cutoff"]),
                 objects=[(g["centers"], g["radii"], None, None)], ws_min=g["ws_min"], ws_max=g["ws_max"])


def test_tiago_sphere_model_oracle_matches_reference():
    g = load_golden("points_tiago_sphere_model")
    d, ids, spheres, build = tiago_with_spheres()
    assert ids == list(g["pt_cols"]) and [s[0] for s in spheres] == list(g["sphere_links"])
    assert np.array_equal(d.trans[ids], g["sphere_offsets"]) and np.array_equal(np.array([s[2] for s in spheres], np.float32), g["pt_radius"])
    cost = tiago_cost(g, ids, spheres, build)
    q = g["q"].reshape(-1, d.n_dofs)
    for o in (orc.objective(d, cost, q, dtype=np.float32), co.objective(co.Tree(d), co.Cost(cost), q)):
        if "pw" in o:
            assert rel_err(o["pw"], g["link_pos"].reshape(q.shape[0], -1, 3)) < 1e-5
        assert rel_err(o["terms"][:, 1], g["cost_obj"].reshape(-1)) < 1e-5
        assert rel_err(o["terms"][:, 2], g["cost_ws"].reshape(-1)) < 1e-5
        assert rel_err(o["grad_q"], g["grad_total"].reshape(-1, d.n_dofs)) < 1e-4
        assert np.mean(((o["occ"] >> 1) & 1) != g["occ_obj"].reshape(-1)) < 0.02


@pytest.mark.gpu
def test_gpu_tiago_sphere_model_matches_reference():
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    g = load_golden("points_tiago_sphere_model")
    d, ids, spheres, build = tiago_with_spheres()
    m = DeviceModel(d, "cuda:0")
    assert m.kernel_family == "generic"
    q = torch.as_tensor(g["q"].reshape(-1, d.n_dofs), device="cuda:0")
    H = m.fk_forward(q)
    assert rel_err(H[:, :, :3, 3].cpu().numpy(), g["link_pos"].reshape(q.shape[0], -1, 3)) < 1e-5
    r = DeviceCost(m, tiago_cost(g, ids, spheres, build)).objective(q, want_terms=True, want_occupancy=True)
    assert rel_err(r["terms"][:, 1].cpu().numpy(), g["cost_obj"].reshape(-1)) < 1e-5
    assert rel_err(r["terms"][:, 2].cpu().numpy(), g["cost_ws"].reshape(-1)) < 1e-5
    assert rel_err(r["grad_q"].cpu().numpy(), g["grad_total"].reshape(-1, d.n_dofs)) < 1e-4
