#!/usr/bin/env python
"""Golden-vector generator: runs the UNMODIFIED reference (/root/reference, CPU, fp32) and records
inputs/outputs of the hot path as small .npz fixtures next to this script.

The reference has no tests or golden vectors of its own, so parity is pinned
by running the reference itself.  It cannot travel to the GPU box, hence the committed fixtures.

Run (in the build container only):
    PYTHONPATH=tests/golden/_shims:/root/reference python tests/golden/make_golden.py

`_shims/` holds stand-ins for `urdf_parser_py` and `matplotlib`, which the reference imports but
which are not installed here.  Nothing in this script is used by the product path.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(1, "/root/reference")

torch.set_num_threads(1)  # deterministic reductions

with contextlib.redirect_stdout(io.StringIO()):
    from torch_robotics.torch_kinematics_tree.models import robots as R
    from torch_robotics.torch_kinematics_tree.geometrics.frame import Frame
    from torch_robotics.torch_kinematics_tree.geometrics.spatial_vector import y_rot, z_rot
    from torch_robotics.torch_kinematics_tree.geometrics.utils import SE3_distance
    from torch_robotics.robots.robot_panda import RobotPanda
    from torch_robotics.environments.env_spheres_3d import EnvSpheres3D
    from torch_robotics.environments.env_base import EnvBase
    from torch_robotics.environments.primitives import MultiSphereField, ObjectField, MultiBoxField, MultiSharpBoxField
    from torch_robotics.tasks.tasks import PlanningTask
    from torch_robotics.torch_planning_objectives.fields.distance_fields import (
        EESE3DistanceField, CollisionObjectDistanceField, CollisionWorkspaceBoundariesDistanceField)

TA = dict(device="cpu", dtype=torch.float32)

ROBOTS = {
    "panda": lambda: R.DifferentiableFrankaPanda(gripper=False),
    "panda_gripper": lambda: R.DifferentiableFrankaPanda(gripper=True),
    "ur10": lambda: R.DifferentiableUR10(),
    "kuka_iiwa": lambda: R.DifferentiableKUKAiiwa(),
    "stretch": lambda: R.DifferentiableHabitatStretch(),
    "tiago_dual_holo": lambda: R.DifferentiableTiagoDualHolo(),
    "tiago_dual_holo_move": lambda: R.DifferentiableTiagoDualHoloMove(),
    "shadow_hand": lambda: R.DifferentiableShadowHand(),
    "allegro_hand": lambda: R.DifferentiableAllegroHand(),
    "planar_2link": lambda: R.Differentiable2LinkPlanar(),
}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def sample_q(model, n, gen):
    """Mix of in-limit, [0,1) (examples/forward_kinematics.py:24), out-of-limit and at-limit rows."""
    D = model._n_dofs
    lo = np.full(D, -np.pi, np.float32)
    hi = np.full(D, np.pi, np.float32)
    for c, bi in enumerate(model._controlled_joints):
        lim = model._bodies[bi].joint_limits
        if lim is not None:
            lo[c], hi[c] = np.float32(lim["lower"]), np.float32(lim["upper"])
    lo_t, hi_t = torch.from_numpy(lo), torch.from_numpy(hi)
    k = n // 4
    q_in = lo_t + torch.rand(k, D, generator=gen) * (hi_t - lo_t)
    q_01 = torch.rand(k, D, generator=gen)
    q_out = (torch.rand(k, D, generator=gen) * 2 - 1) * 3.9
    q_at = torch.where(torch.rand(n - 3 * k, D, generator=gen) < 0.5, lo_t, hi_t)
    return torch.cat([q_in, q_01, q_out, q_at], 0).contiguous()


def dump_descriptor(model):
    bodies = model._bodies
    names = [b.name for b in bodies]
    idx = {n: i for i, n in enumerate(names)}
    parent = [-1] + [idx[model._model.get_name_of_parent_body(b.name)] for b in bodies[1:]]
    lo, hi, has = [], [], []
    for b in bodies:
        lim = b.joint_limits
        has.append(0 if lim is None else 1)
        lo.append(0.0 if lim is None else lim["lower"])
        hi.append(0.0 if lim is None else lim["upper"])
    return dict(
        link_names=np.array(names), parent=np.array(parent, np.int32),
        joint_type=np.array([b.joint_type for b in bodies]),
        fixed_rotation=torch.cat([b.fixed_rotation for b in bodies]).numpy(),
        trans=torch.cat([b.trans for b in bodies]).numpy(),
        joint_axis=torch.cat([b.joint_axis for b in bodies]).numpy(),
        joint_idx=np.array([-1 if b.joint_idx is None else b.joint_idx for b in bodies], np.int32),
        limit_lower=np.array(lo, np.float64), limit_upper=np.array(hi, np.float64),
        has_limit_dict=np.array(has, np.int32),
        out_link_names=np.array(model.get_link_names()),
    )


def golden_fk():
    gen = torch.Generator().manual_seed(1234)
    for name, ctor in ROBOTS.items():
        model = quiet(ctor)
        out = dump_descriptor(model)
        q = sample_q(model, 32, gen).requires_grad_(True)
        H = model.compute_forward_kinematics_all_links(q)
        G = torch.randn(H.shape, generator=gen)
        (H * G).sum().backward()
        out.update(q=q.detach().numpy(), H=H.detach().numpy(), G=G.numpy(), grad_q=q.grad.numpy())
        np.savez_compressed(os.path.join(HERE, f"fk_{name}.npz"), **out)
        print(name, "links", len(model._bodies), "dofs", model._n_dofs, "H", tuple(H.shape))


def ee_target():
    # target of examples/inverse_kinematics.py:27-30
    pos = torch.tensor([0.2, 0.4, 0.1], **TA)
    rot = (z_rot(-torch.tensor(torch.pi / 2)) @ y_rot(-torch.tensor(torch.pi))).to(**TA)
    return Frame(rot=rot, trans=pos, device="cpu").get_transform_matrix()   # [1,4,4]


def obstacle_argmin(env, pos):
    """Index of the nearest sphere per collision point, from the reference's own primitive tensors."""
    out = []
    for obj in env.get_df_obj_list():
        if len(obj.fields) != 1 or not isinstance(obj.fields[0], MultiSphereField):
            out.append(torch.full(pos.shape[:-1], -1, dtype=torch.int64))     # arg-min recorded for sphere-only objects
            continue
        f = obj.fields[0]
        from torch_robotics.torch_kinematics_tree.geometrics.quaternion import q_to_rotation_matrix
        from torch_robotics.torch_kinematics_tree.geometrics.utils import rotate_point
        rot = q_to_rotation_matrix(obj.ori).squeeze().transpose(-2, -1)
        x = rotate_point(pos - obj.pos, rot)
        sdfs = torch.norm(x.unsqueeze(-2) - f.centers.unsqueeze(0), dim=-1) - f.radii
        out.append(torch.min(sdfs, dim=-1)[1])
    return torch.stack(out, dim=-2)   # [..., n_obj, S]


def planning_case(tag, robot, env, task, q, H_target, ee_kwargs, ws=True):
    """Record every term of the objective and its gradient for q [b,h,7]."""
    out = {}
    b, h, _ = q.shape
    terms = {}

    def run(which):
        qq = q.clone().requires_grad_(True)
        pos = robot.fk_map_collision(qq)
        if which == "self":
            c = task.df_collision_self.compute_cost(qq, pos, field_type="sdf")
        elif which == "obj":
            c = task.df_collision_objects.compute_cost(qq, pos, field_type="sdf")
        elif which == "ws":
            c = task.df_collision_ws_boundaries.compute_cost(qq, pos, field_type="sdf")
        elif which == "ee":
            Hs = robot.diff_panda.compute_forward_kinematics_all_links(qq.reshape(-1, 7))
            ee = EESE3DistanceField(H_target, **ee_kwargs)
            c = ee.compute_costs_impl(qq, Hs).reshape(b, h)
        elif which == "collision":
            c = task.compute_collision_cost(qq)
        c.sum().backward()
        return c.detach().numpy(), qq.grad.numpy(), pos.detach().numpy()

    for which in ("self", "obj", "ws", "ee", "collision"):
        c, g, pos = run(which)
        out[f"cost_{which}"] = c
        out[f"grad_{which}"] = g
    out["link_pos"] = pos
    out["q"] = q.numpy()
    out["H_target"] = H_target.numpy().reshape(4, 4)
    # occupancy masks: per field, default margins and margin=0 (tasks.py:251)
    with torch.no_grad():
        pos_t = robot.fk_map_collision(q)
        for mtag, kw in (("default", {}), ("zero", {"margin": 0.0})):
            out[f"occ_self_{mtag}"] = task.df_collision_self.compute_cost(q, pos_t, field_type="occupancy", **kw).numpy()
            out[f"occ_obj_{mtag}"] = task.df_collision_objects.compute_cost(q, pos_t, field_type="occupancy", **kw).numpy()
            out[f"occ_ws_{mtag}"] = task.df_collision_ws_boundaries.compute_cost(q, pos_t, field_type="occupancy", **kw).numpy()
            out[f"occ_all_{mtag}"] = task.compute_collision(q, **kw).numpy()
        pts = pos_t[..., robot.link_idxs_for_object_collision_checking, :]
        out["obj_argmin"] = obstacle_argmin(env, pts).numpy().astype(np.int32)
    # environment description, so tests do not need the reference classes
    objs = env.get_df_obj_list()
    out["n_obj"] = np.int32(len(objs))
    for oi, obj in enumerate(objs):
        sph = [f for f in obj.fields if isinstance(f, MultiSphereField)]
        out[f"obj{oi}_centers"] = torch.cat([f.centers for f in sph]).numpy() if sph else np.zeros((0, 3), np.float32)
        out[f"obj{oi}_radii"] = torch.cat([f.radii for f in sph]).numpy() if sph else np.zeros(0, np.float32)
        out[f"obj{oi}_pos"] = obj.pos.numpy()
        out[f"obj{oi}_ori_wxyz"] = obj.ori.numpy()
        boxes = [f for f in obj.fields if isinstance(f, MultiSharpBoxField)]
        out[f"obj{oi}_n_box_fields"] = np.int32(len(boxes))
        for bi, f in enumerate(boxes):
            out[f"obj{oi}_box{bi}_centers"] = f.centers.numpy()
            out[f"obj{oi}_box{bi}_sizes"] = f.sizes.numpy()
            out[f"obj{oi}_box{bi}_rounded"] = np.int32(isinstance(f, MultiBoxField))
    out["ws_min"] = task.ws_min.numpy()
    out["ws_max"] = task.ws_max.numpy()
    out["pt_link"] = np.array(robot.link_idxs_for_object_collision_checking, np.int32)
    out["pt_radius"] = robot.link_margins_for_object_collision_checking_tensor.numpy()
    out["obstacle_cutoff_margin"] = np.float64(task.df_collision_objects.cutoff_margin)
    out["self_link_idxs"] = np.array(robot.link_idxs_for_self_collision_checking, np.int32)
    out["self_pairs_local"] = np.array(robot.df_collision_self.idxs_links_distance_matrix, np.int32)
    out["self_margin"] = robot.df_collision_self.cutoff_margin.numpy()
    out["ee_w_pos"] = np.float64(ee_kwargs.get("w_pos", 1.0))
    out["ee_w_rot"] = np.float64(ee_kwargs.get("w_rot", 1.0))
    out["ee_square"] = np.int32(ee_kwargs.get("square", True))
    np.savez_compressed(os.path.join(HERE, f"planning_{tag}.npz"), **out)
    print("planning", tag, "cost_collision[0,:3]", out["cost_collision"][0, :3])


def golden_planning():
    gen = torch.Generator().manual_seed(4321)
    robot = quiet(RobotPanda, tensor_args=TA)
    H_target = ee_target()

    def sample(b, h):
        lo, hi = robot.q_min, robot.q_max
        q = lo + torch.rand(b, h, 7, generator=gen) * (hi - lo)
        q[0, 0] = torch.rand(7, generator=gen) * 7.8 - 3.9     # out-of-limit row
        return q.contiguous()

    # (1) config 3 of BASELINE.json: EnvSpheres3D, default PlanningTask
    env = quiet(EnvSpheres3D, tensor_args=TA)
    task = quiet(PlanningTask, env=env, robot=robot, tensor_args=TA)
    planning_case("panda_spheres3d", robot, env, task, sample(8, 16), H_target, dict(w_pos=1.0, w_rot=1.0, square=True))

    # (2) two objects, one with a non-identity pose, non-uniform radii, un-squared EE weights
    c1 = torch.tensor([[0.4, 0.1, 0.5], [-0.2, 0.4, 0.6], [0.1, -0.5, 0.3]])
    r1 = torch.tensor([0.12, 0.2, 0.07])
    c2 = torch.tensor([[0.0, 0.0, 0.0], [0.25, 0.0, 0.1], [0.0, 0.3, -0.1], [-0.3, -0.2, 0.2]])
    r2 = torch.tensor([0.1, 0.15, 0.05, 0.2])
    o1 = ObjectField([MultiSphereField(c1, r1, tensor_args=TA)], "o1")
    quat = torch.tensor([0.8, 0.2, -0.4, 0.4], **TA)
    quat = quat / quat.norm()
    o2 = ObjectField([MultiSphereField(c2, r2, tensor_args=TA)], "o2",
                     pos=torch.tensor([0.3, -0.1, 0.5], **TA), ori=quat)
    env2 = EnvBase(name="two_objects", limits=torch.tensor([[-0.6, -0.7, -0.1], [0.7, 0.6, 1.0]], **TA),
                   obj_fixed_list=[o1, o2], tensor_args=TA)
    task2 = quiet(PlanningTask, env=env2, robot=robot, obstacle_cutoff_margin=0.03, tensor_args=TA)
    planning_case("panda_two_objects", robot, env2, task2, sample(4, 8), H_target, dict(w_pos=2.0, w_rot=0.5, square=False))


def golden_planning_boxes():
    """Rounded + sharp boxes (primitives.py:198-334), mixed with spheres, one object posed: the table/shelf style
    environments of the reference."""
    gen = torch.Generator().manual_seed(808)
    robot = quiet(RobotPanda, tensor_args=TA)
    H_target = ee_target()
    rounded = MultiBoxField(torch.tensor([[0.45, 0.0, 0.25], [-0.3, 0.4, 0.5]]), torch.tensor([[0.3, 0.6, 0.5], [0.25, 0.2, 0.4]]), tensor_args=TA)
    balls = MultiSphereField(torch.tensor([[0.1, -0.45, 0.6], [0.3, 0.35, 0.8]]), torch.tensor([0.12, 0.1]), tensor_args=TA)
    o1 = ObjectField([balls, rounded], "table_like")
    sharp = MultiSharpBoxField(torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.25, 0.3]]), torch.tensor([[0.5, 0.05, 0.6], [0.5, 0.3, 0.04]]), tensor_args=TA)
    quat = torch.tensor([0.9, 0.0, 0.1, 0.4], **TA)
    quat = quat / quat.norm()
    o2 = ObjectField([sharp], "shelf_like", pos=torch.tensor([-0.45, -0.35, 0.4], **TA), ori=quat)
    env = EnvBase(name="boxes", limits=torch.tensor([[-1, -1, -1], [1, 1, 1]], **TA), obj_fixed_list=[o1, o2], tensor_args=TA)
    task = quiet(PlanningTask, env=env, robot=robot, tensor_args=TA)
    lo, hi = robot.q_min, robot.q_max
    q = (lo + torch.rand(8, 16, 7, generator=gen) * (hi - lo)).contiguous()
    planning_case("panda_boxes", robot, env, task, q, H_target, dict(w_pos=1.0, w_rot=1.0, square=True))


def golden_planning_envs():
    """The reference's other Panda 3-D environments (SURVEY.md section 8 row f2 names them): EnvTableShelf (a posed
    table + a posed ten-board shelf, rounded boxes), EnvMazeBoxes3D (fourteen rounded pillars) and
    EnvSpheres3DExtraObjects (fixed + extra sphere objects)."""
    from torch_robotics.environments.env_table_shelf import EnvTableShelf
    from torch_robotics.environments.env_maze_boxes_3d import EnvMazeBoxes3D
    from torch_robotics.environments.env_spheres_3d_extra_objects import EnvSpheres3DExtraObjects
    gen = torch.Generator().manual_seed(1123)
    robot = quiet(RobotPanda, tensor_args=TA)
    H_target = ee_target()
    for tag, cls in (("table_shelf", EnvTableShelf), ("maze_boxes", EnvMazeBoxes3D), ("spheres3d_extra", EnvSpheres3DExtraObjects)):
        env = quiet(cls, tensor_args=TA)
        task = quiet(PlanningTask, env=env, robot=robot, tensor_args=TA)
        q = (robot.q_min + torch.rand(6, 12, 7, generator=gen) * (robot.q_max - robot.q_min)).contiguous()
        planning_case(f"panda_{tag}", robot, env, task, q, H_target, dict(w_pos=1.0, w_rot=1.0, square=True))
        g = dict(np.load(os.path.join(HERE, f"planning_panda_{tag}.npz")))
        g["env_limits"] = env.limits.numpy()
        np.savez_compressed(os.path.join(HERE, f"planning_panda_{tag}.npz"), **g)


def golden_stretch_objective():
    """Config 5: Stretch has no collision model in the reference; compose the generic field classes
    exactly as SURVEY.md section 8(d) item 5 specifies."""
    gen = torch.Generator().manual_seed(99)
    model = quiet(R.DifferentiableHabitatStretch)
    names = model.get_link_names()
    pt_links = list(model._controlled_joints) + [names.index("link_straight_gripper")]
    radius = torch.full((len(pt_links),), 0.05, **TA)
    env = quiet(EnvSpheres3D, tensor_args=TA)

    class _Stub:   # the fields only read robot.grasped_object (distance_fields.py:137)
        grasped_object = None
    f_obj = CollisionObjectDistanceField(_Stub(), df_obj_list_fn=env.get_df_obj_list,
                                         link_idxs_for_collision_checking=pt_links,
                                         link_margins_for_object_collision_checking_tensor=radius,
                                         cutoff_margin=0.01, tensor_args=TA)
    f_ws = CollisionWorkspaceBoundariesDistanceField(_Stub(), link_idxs_for_collision_checking=pt_links,
                                                     link_margins_for_object_collision_checking_tensor=radius,
                                                     cutoff_margin=0.01, ws_bounds_min=env.limits[0],
                                                     ws_bounds_max=env.limits[1], tensor_args=TA)
    H_target = ee_target()
    ee_idx = names.index("link_straight_gripper")
    q = sample_q(model, 32, gen)
    out = dict(q=q.numpy(), pt_link=np.array(pt_links, np.int32), pt_radius=radius.numpy(), ee_link=np.int32(ee_idx),
               H_target=H_target.numpy().reshape(4, 4), centers=env.obj_fixed_list[0].fields[0].centers.numpy(),
               radii=env.obj_fixed_list[0].fields[0].radii.numpy(), ws_min=env.limits[0].numpy(), ws_max=env.limits[1].numpy())
    for which in ("obj", "ws", "ee", "total"):
        qq = q.clone().requires_grad_(True)
        Hs = model.compute_forward_kinematics_all_links(qq)
        pos = Hs[..., :3, 3]
        c_obj = f_obj.compute_cost(qq, pos, field_type="sdf").reshape(-1)
        c_ws = f_ws.compute_cost(qq, pos, field_type="sdf").reshape(-1)
        c_ee = torch.square(SE3_distance(Hs[:, ee_idx], H_target, w_pos=1.0, w_rot=1.0))
        c = dict(obj=c_obj, ws=c_ws, ee=c_ee, total=c_obj + c_ws + c_ee)[which]
        c.sum().backward()
        out[f"cost_{which}"] = c.detach().numpy()
        out[f"grad_{which}"] = qq.grad.numpy()
    with torch.no_grad():
        pos = model.compute_forward_kinematics_all_links(q)[..., :3, 3]
        out["occ_obj"] = f_obj.compute_cost(q, pos, field_type="occupancy").reshape(-1).numpy()
        out["occ_ws"] = f_ws.compute_cost(q, pos, field_type="occupancy").reshape(-1).numpy()
    np.savez_compressed(os.path.join(HERE, "objective_stretch.npz"), **out)
    print("stretch objective total[:3]", out["cost_total"][:3])


def golden_ik():
    gen = torch.Generator().manual_seed(7)
    model = quiet(R.DifferentiableFrankaPanda, gripper=False)
    H_target = ee_target()
    lower, upper, _, _ = model.get_joint_limit_array()
    eps = torch.pi / 64
    lo = torch.tensor(lower + eps, **TA)
    hi = torch.tensor(upper - eps, **TA)
    S = 24
    q0 = (lo + torch.rand(S, 7, generator=gen) * (hi - lo)).contiguous()
    q_probe = q0.clone()
    q_probe[:4] += 0.3    # some rows outside the shrunk limits
    qq = q_probe.clone().requires_grad_(True)
    err = model.loss_fn_ik_per_q(qq, H_target, "ee_link", w_se3=1.0, w_joint_limits=300.0, lower=lo, upper=hi,
                                 w_q_rest=1.0, q_rest=None)
    err.sum().backward()
    idx_valid = model.ik_termination(q_probe, H_target, "ee_link", lo, hi, se3_eps=0.5)
    out = dict(q_probe=q_probe.numpy(), loss=err.detach().numpy(), loss_grad=qq.grad.numpy(),
               idx_valid_eps0p5=idx_valid.numpy().astype(np.int64), lower=lo.numpy(), upper=hi.numpy(),
               H_target=H_target.numpy().reshape(4, 4), q0=q0.numpy())
    for iters in (1, 5, 40):
        q_ik, idx = quiet(model.inverse_kinematics, H_target, link_name="ee_link", batch_size=S, max_iters=iters,
                          lr=2e-1, se3_eps=5e-2, q0=q0.clone(), q0_noise=0.0, eps_joint_lim=eps, print_freq=-1)
        out[f"q_ik_{iters}"] = q_ik.detach().numpy()
        out[f"idx_valid_{iters}"] = idx.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "ik_panda.npz"), **out)
    print("ik: valid after 40 iters:", out["idx_valid_40"])




def golden_ik_500():
    """BASELINE config 4's settings (examples/inverse_kinematics.py:37-42: 500 iterations, lr 0.2, se3_eps 0.05) on 64
    explicit seeds.  Adam's normalised steps make the iterates chaotic over hundreds of iterations: two fp32
    implementations of the same loop drift apart (the numpy restatement vs this reference run: median 0.06 rad, three of 64
    validity flags differ), so the fixture also records the per-seed SE(3) error for the tests' boundary analysis."""
    gen = torch.Generator().manual_seed(77)
    model = quiet(R.DifferentiableFrankaPanda, gripper=False)
    H_target = ee_target()
    lower, upper, _, _ = model.get_joint_limit_array()
    eps = torch.pi / 64
    lo = torch.tensor(lower + eps, **TA)
    hi = torch.tensor(upper - eps, **TA)
    S = 64
    q0 = (lo + torch.rand(S, 7, generator=gen) * (hi - lo)).contiguous()
    q_ik, idx = quiet(model.inverse_kinematics, H_target, link_name="ee_link", batch_size=S, max_iters=500, lr=2e-1, se3_eps=5e-2,
                      q0=q0.clone(), q0_noise=0.0, eps_joint_lim=eps, print_freq=-1)
    H = model.compute_forward_kinematics_all_links(q_ik.detach(), link_list=["ee_link"]).squeeze(1)
    err = SE3_distance(H, H_target).detach().numpy().reshape(-1)
    np.savez_compressed(os.path.join(HERE, "ik_panda_500.npz"), q0=q0.numpy(), q_ik_500=q_ik.detach().numpy(),
                        idx_valid_500=idx.numpy().astype(np.int64), err_500=err.astype(np.float32), lower=lo.numpy(), upper=hi.numpy(),
                        H_target=H_target.numpy().reshape(4, 4))
    print("ik 500: valid", len(idx), "of", S)


def golden_trajectories():
    """Trajectory collision classification (tasks.py:234-328) on Panda + EnvSpheres3D: interpolation between
    waypoints, zero-margin collision flags, free / colliding index lists, joint-limit filter, statistics."""
    from torch_robotics.trajectory.utils import interpolate_traj_via_points
    gen = torch.Generator().manual_seed(2024)
    robot = quiet(RobotPanda, tensor_args=TA)
    env = quiet(EnvSpheres3D, tensor_args=TA)
    task = quiet(PlanningTask, env=env, robot=robot, tensor_args=TA)
    lo, hi = robot.q_min, robot.q_max
    B, H = 24, 12
    # short straight-line segments around random configurations: a mix of free and colliding trajectories
    start = lo + torch.rand(B, 1, 7, generator=gen) * (hi - lo)
    step = (torch.rand(B, 1, 7, generator=gen) - 0.5) * 0.04
    trajs = (start + step * torch.arange(H).view(1, H, 1)).contiguous()
    trajs[3, 5, 0] = hi[0] + 0.2          # leaves the joint limits
    out = {"trajs": trajs.numpy()}
    out["interp5"] = interpolate_traj_via_points(trajs, num_interpolation=5).numpy()
    trajs_in_limits = torch.maximum(torch.minimum(trajs, hi), lo).contiguous()
    out["trajs_in_limits"] = trajs_in_limits.numpy()
    # the 4-D input path of the reference raises when a collision-free trajectory violates the joint limits
    # (torch.cat of [n,2] and [m,1] index tensors, tasks.py:281-283), so it is recorded without such a trajectory
    for tag, t in (("3d", trajs), ("4d", trajs_in_limits.reshape(2, 12, H, 7))):
        coll, coll_idx, free, free_idx, wp = task.get_trajs_collision_and_free(t, return_indices=True, num_interpolation=5)
        out[f"{tag}_coll_idx"] = coll_idx.numpy().astype(np.int64)
        out[f"{tag}_free_idx"] = free_idx.numpy().astype(np.int64)
        out[f"{tag}_waypoint_collisions"] = wp.numpy()
        out[f"{tag}_n_coll"] = np.int64(0 if coll is None else coll.shape[-3])
        out[f"{tag}_n_free"] = np.int64(0 if free is None else free.shape[-3])
        out[f"{tag}_coll"] = np.zeros(0, np.float32) if coll is None else coll.numpy()
        out[f"{tag}_free"] = np.zeros(0, np.float32) if free is None else free.numpy()
    out["fraction_free"] = np.float64(task.compute_fraction_free_trajs(trajs))
    out["collision_intensity"] = np.float64(task.compute_collision_intensity_trajs(trajs))
    out["success"] = np.int64(task.compute_success_free_trajs(trajs))
    np.savez_compressed(os.path.join(HERE, "planning_trajectories.npz"), **out)
    print("trajectories: free", out["3d_free_idx"].reshape(-1), "coll", out["3d_coll_idx"].reshape(-1), "frac", out["fraction_free"])


class _NoGrasp:   # the generic fields only read robot.grasped_object (distance_fields.py:137)
    grasped_object = None


def _point_model_case(tag, q, pos_fn, pt_cols, pt_radius, env, cutoff, self_rows=None, self_pairs=None, self_margin=None, extra=None):
    """Objective over collision points that are columns of pos_fn(q) [..., C, 3], composed from the reference's
    generic field classes (the pattern of golden_stretch_objective)."""
    f_obj = CollisionObjectDistanceField(_NoGrasp(), df_obj_list_fn=env.get_df_obj_list,
                                         link_idxs_for_collision_checking=pt_cols,
                                         link_margins_for_object_collision_checking_tensor=pt_radius,
                                         cutoff_margin=cutoff, tensor_args=TA)
    f_ws = CollisionWorkspaceBoundariesDistanceField(_NoGrasp(), link_idxs_for_collision_checking=pt_cols,
                                                     link_margins_for_object_collision_checking_tensor=pt_radius,
                                                     cutoff_margin=cutoff, ws_bounds_min=env.limits[0],
                                                     ws_bounds_max=env.limits[1], tensor_args=TA)
    f_self = None
    if self_rows is not None:
        from torch_robotics.torch_planning_objectives.fields.distance_fields import CollisionSelfField
        f_self = CollisionSelfField(_NoGrasp(), link_idxs_for_collision_checking=self_rows,
                                    idxs_links_distance_matrix=self_pairs, cutoff_margin=self_margin, tensor_args=TA)
    fields = dict(obj=f_obj, ws=f_ws)
    if f_self is not None:
        fields["self"] = f_self
    out = dict(q=q.numpy(), pt_cols=np.array(pt_cols, np.int32), pt_radius=pt_radius.numpy(), cutoff=np.float64(cutoff),
               centers=env.obj_fixed_list[0].fields[0].centers.numpy(), radii=env.obj_fixed_list[0].fields[0].radii.numpy(),
               ws_min=env.limits[0].numpy(), ws_max=env.limits[1].numpy())
    if f_self is not None:
        out.update(self_rows=np.array(self_rows, np.int32), self_pairs_local=np.array(self_pairs, np.int32),
                   self_margin=self_margin.numpy())
    out.update(extra or {})
    for which in list(fields) + ["total"]:
        qq = q.clone().requires_grad_(True)
        pos = pos_fn(qq)
        cs = {k: f.compute_cost(qq, pos, field_type="sdf") for k, f in fields.items()}
        c = sum(cs.values()) if which == "total" else cs[which]
        c.sum().backward()
        out[f"cost_{which}"] = c.detach().numpy()
        out[f"grad_{which}"] = qq.grad.numpy()
    with torch.no_grad():
        pos = pos_fn(q)
        out["link_pos"] = pos.numpy()
        for k, f in fields.items():
            out[f"occ_{k}"] = f.compute_cost(q, pos, field_type="occupancy").numpy()
    np.savez_compressed(os.path.join(HERE, f"points_{tag}.npz"), **out)
    print("points", tag, "cost_total[0,:3]", out["cost_total"].reshape(-1)[:3], "occ", {k: int(out[f"occ_{k}"].sum()) for k in fields})


def golden_attached_points():
    """collision points that are not link origins.
    (1) the grasped box of environments/objects.py:37-89 on RobotPanda.  The reference's PlanningTask raises
        IndexError for it (distance_fields.py:137-154 selects the link columns twice), so the objective is composed
        from its generic field classes over the columns robot_base.py:75-127 sets up: the five link spheres + the 14
        box points against the obstacles / workspace, the ten robot pairs + 14 x 4 box-link pairs for self-collision.
    (2) the per-link multi-sphere model of data/configs/panda/panda_sphere_config.yaml (unused by the reference's
        own code): sphere centres are Frame.transform_point() of the link frames."""
    import yaml
    from pathlib import Path
    from torch_robotics.environments.objects import GraspedObjectPandaBox
    gen = torch.Generator().manual_seed(2718)
    env = quiet(EnvSpheres3D, tensor_args=TA)

    # (1) the reference rewrites the URDF next to the original; the tree is read-only here and already holds the
    # result of that rewriting for this very object, so hand its path back instead of writing.
    R.modidy_franka_panda_urdf_grasped_object = lambda f, g: Path(str(f).replace(".urdf", "_grasped_object.urdf"))
    go = GraspedObjectPandaBox(tensor_args=TA)
    robot = quiet(RobotPanda, grasped_object=go, tensor_args=TA)
    n_links = len(robot.diff_panda.get_link_names())
    n_pts = go.n_base_points_for_collision
    grasp_cols = list(range(n_links, n_links + n_pts))
    q = (robot.q_min + torch.rand(16, 8, 7, generator=gen) * (robot.q_max - robot.q_min)).contiguous()
    _point_model_case("panda_grasped_box", q, robot.fk_map_collision,
                      list(robot.link_idxs_for_object_collision_checking) + grasp_cols,
                      robot.link_margins_for_object_collision_checking_tensor, env, 0.02,
                      self_rows=list(robot.link_idxs_for_self_collision_checking) + grasp_cols,
                      self_pairs=robot.df_collision_self.idxs_links_distance_matrix,
                      self_margin=robot.df_collision_self.cutoff_margin,
                      extra=dict(base_points=go.base_points_for_collision.numpy(), grasp_pos=go.pos.numpy(), grasp_ori=go.ori.numpy(),
                                 link_names=np.array(robot.diff_panda.get_link_names())))

    # (2)
    cfg = yaml.safe_load(open("/root/reference/torch_robotics/data/configs/panda/panda_sphere_config.yaml"))
    robot = quiet(RobotPanda, tensor_args=TA)
    names = robot.diff_panda.get_link_names()
    spheres = [(l, s) for l, lst in cfg.items() if l in names for s in lst]
    offs = torch.tensor([s[:3] for _, s in spheres], **TA)
    radii = torch.tensor([s[3] for _, s in spheres], **TA)

    def pos_fn(qq):
        shp = qq.shape
        frames = robot.diff_panda.compute_forward_kinematics_all_links(qq.reshape(-1, 7), return_dict=True)
        link_pos = torch.stack([frames[n].translation for n in names], dim=-2)
        pts = torch.stack([frames[l].transform_point(offs[i:i + 1]).reshape(-1, 3) for i, (l, _) in enumerate(spheres)], dim=-2)
        return torch.cat((link_pos, pts), dim=-2).reshape(*shp[:-1], len(names) + len(spheres), 3)

    q = (robot.q_min + torch.rand(12, 8, 7, generator=gen) * (robot.q_max - robot.q_min)).contiguous()
    _point_model_case("panda_sphere_model", q, pos_fn, list(range(len(names), len(names) + len(spheres))), radii, env, 0.01,
                      extra=dict(sphere_links=np.array([l for l, _ in spheres]), sphere_offsets=offs.numpy()))

    # (3) the Tiago sphere model (data/configs/tiago/tiago_sphere_config.yaml: [x, y, z, radius, group] per link) on the
    # branching dual-arm holonomic-base robot: sphere centres and their Jacobian-free gradient path through a 19-link tree
    cfg = yaml.safe_load(open("/root/reference/torch_robotics/data/configs/tiago/tiago_sphere_config.yaml"))
    model = quiet(ROBOTS["tiago_dual_holo"])
    tnames = model.get_link_names()
    tsph = [(l, s) for l, lst in cfg.items() if isinstance(lst, list) and l in tnames for s in lst]
    toffs = torch.tensor([s[:3] for _, s in tsph], **TA)
    tradii = torch.tensor([s[3] for _, s in tsph], **TA)

    def tiago_pos(qq):
        shp = qq.shape
        frames = model.compute_forward_kinematics_all_links(qq.reshape(-1, shp[-1]), return_dict=True)
        link_pos = torch.stack([frames[n].translation for n in tnames], dim=-2)
        pts = torch.stack([frames[l].transform_point(toffs[i:i + 1]).reshape(-1, 3) for i, (l, _) in enumerate(tsph)], dim=-2)
        return torch.cat((link_pos, pts), dim=-2).reshape(*shp[:-1], len(tnames) + len(tsph), 3)

    qt = sample_q(model, 64, gen).reshape(8, 8, -1).contiguous()
    _point_model_case("tiago_sphere_model", qt, tiago_pos, list(range(len(tnames), len(tnames) + len(tsph))), tradii, env, 0.01,
                      extra=dict(sphere_links=np.array([l for l, _ in tsph]), sphere_offsets=toffs.numpy()))


def golden_voxel_sdf():
    """EnvBase(precompute_sdf_obj_fixed=True) replaces the fixed objects by a GridMapSDF
    (grid_map_sdf.py:9-114).  Coarse cells (0.05 m -> 40^3) keep the fixture small; one extra (non-grid) object is
    added so that the max over objects (distance_fields.py:121) sees both kinds."""
    gen = torch.Generator().manual_seed(1618)
    robot = quiet(RobotPanda, tensor_args=TA)
    extra = ObjectField([MultiSphereField(torch.tensor([[0.3, 0.3, 0.6], [-0.2, -0.4, 0.7]]), torch.tensor([0.2, 0.15]), tensor_args=TA)], "extra")
    env = quiet(EnvSpheres3D, precompute_sdf_obj_fixed=True, sdf_cell_size=0.05, obj_extra_list=[extra], tensor_args=TA)
    grid = env.grid_map_sdf_obj_fixed
    task = quiet(PlanningTask, env=env, robot=robot, tensor_args=TA)
    q = (robot.q_min + torch.rand(8, 16, 7, generator=gen) * (robot.q_max - robot.q_min)).contiguous()
    out = dict(q=q.numpy(), sdf_tensor=grid.sdf_tensor.numpy(), grad_sdf_tensor=grid.grad_sdf_tensor.numpy(),
               cmap_dim=grid.cmap_dim.numpy().astype(np.int32), limits=env.limits.numpy(), cell_size=np.float64(0.05),
               extra_centers=extra.fields[0].centers.numpy(), extra_radii=extra.fields[0].radii.numpy(),
               pt_link=np.array(robot.link_idxs_for_object_collision_checking, np.int32),
               pt_radius=robot.link_margins_for_object_collision_checking_tensor.numpy(),
               obstacle_cutoff_margin=np.float64(task.df_collision_objects.cutoff_margin))
    for which in ("obj", "collision"):
        qq = q.clone().requires_grad_(True)
        pos = robot.fk_map_collision(qq)
        c = task.df_collision_objects.compute_cost(qq, pos, field_type="sdf") if which == "obj" else task.compute_collision_cost(qq)
        c.sum().backward()
        out[f"cost_{which}"], out[f"grad_{which}"] = c.detach().numpy(), qq.grad.numpy()
    with torch.no_grad():
        pos = robot.fk_map_collision(q)
        out["link_pos"] = pos.numpy()
        out["occ_obj"] = task.df_collision_objects.compute_cost(q, pos, field_type="occupancy").numpy()
        out["occ_obj_zero"] = task.df_collision_objects.compute_cost(q, pos, field_type="occupancy", margin=0.0).numpy()
        out["occ_all"] = task.compute_collision(q).numpy()
        pts = pos[..., robot.link_idxs_for_object_collision_checking, :]
        cells = ((pts - grid.limits[0]) / grid.map_dim * grid.cmap_dim).floor().to(torch.int)
        hi = torch.tensor(grid.points_for_sdf.shape[:-1]).type_as(cells) - 1
        out["cells"] = cells.clamp(torch.zeros_like(hi), hi).numpy()
        # standalone queries incl. points outside the workspace (clamped to the border cells, grid_map_sdf.py:99-100)
        X = (torch.rand(512, 3, generator=gen) * 2.6 - 1.3).requires_grad_(False)
        out["query_x"] = X.numpy()
        out["query_sdf"] = grid(X.clone()).numpy()
    Xg = torch.tensor(out["query_x"]).requires_grad_(True)
    grid(Xg).sum().backward()
    out["query_grad"] = Xg.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "planning_panda_voxel_sdf.npz"), **out)
    print("voxel sdf", tuple(out["cmap_dim"]), "cost_collision[0,:3]", out["cost_collision"][0, :3], "occ", int(out["occ_obj"].sum()))


def golden_link_jacobians():
    """compute_analytical_jacobian_all_links (robot_tree.py:250-265) for a chain (Panda), a tree with prismatic
    joints and many fixed links (Stretch) and a branching hand with an oblique axis (Shadow); rows cover in-limit,
    [0,1), out-of-limit (clamped -> zero columns) and at-limit configurations."""
    gen = torch.Generator().manual_seed(577)
    out = {}
    for name in ("panda", "stretch", "shadow_hand"):
        model = quiet(ROBOTS[name])
        q = sample_q(model, 8, gen)
        J = model.compute_analytical_jacobian_all_links(q)
        H = model.compute_forward_kinematics_all_links(q)
        out[f"{name}_q"], out[f"{name}_J"], out[f"{name}_H"] = q.numpy(), J.numpy(), H.numpy()
        print("jacobian", name, tuple(J.shape), float(J.abs().max()))
    np.savez_compressed(os.path.join(HERE, "link_jacobians.npz"), **out)


def golden_geometric_jacobian():
    """compute_forward_kinematics_and_geometric_jacobian (robot_tree.py:218-248) for links of the Panda (chain, z axes), the
    UR10 (y and z axes) and the dual-arm Tiago (branches: the reference's joint-index rule admits joints that are not
    ancestors of the link).  These robots have only positive-axis, limited revolute joints: for them the stateful pose
    propagation the reference uses here (update_kinematic_state) gives the same poses as its stateless FK.  (It does not for
    negative axes - rigid_body.py:101-106 picks z_rot and drops the sign - nor for continuous joints, which it clamps.)"""
    gen = torch.Generator().manual_seed(99)
    out = {}
    for name, links in (("panda", ["ee_link", "panda_link4"]), ("ur10", ["ee_link", "forearm_link"]),
                        ("tiago_dual_holo", ["arm_right_7_link", "arm_left_4_link"])):
        model = quiet(ROBOTS[name])
        names = model.get_link_names()
        links = [ln if ln in names else names[-1] for ln in links]
        q = sample_q(model, 24, gen)
        out[f"{name}_q"] = q.numpy()
        out[f"{name}_links"] = np.array(links)
        # stateless FK first: the stateful call below overwrites body.pose, which the stateless recursion starts from
        Hs_all = {ln: model.compute_forward_kinematics_all_links(q, link_list=[ln])[:, 0].clone() for ln in links}
        for ln in links:
            pos, rot, lin, ang = model.compute_forward_kinematics_and_geometric_jacobian(q, torch.zeros_like(q), ln)
            assert torch.allclose(Hs_all[ln][:, :3, 3], pos, atol=1e-6), (name, ln)      # stateful == stateless for these robots
            out[f"{name}_{ln}_pos"], out[f"{name}_{ln}_rot"] = pos.detach().numpy(), rot.detach().numpy()
            out[f"{name}_{ln}_lin"], out[f"{name}_{ln}_ang"] = lin.detach().numpy(), ang.detach().numpy()
            print("geometric jacobian", name, ln, tuple(lin.shape), float(lin.abs().max()))
    np.savez_compressed(os.path.join(HERE, "geometric_jacobian.npz"), **out)


if __name__ == "__main__":
    if os.environ.get("GOLDEN_ONLY") == "envs":
        golden_planning_envs()
    elif os.environ.get("GOLDEN_ONLY") == "jacobian":
        golden_link_jacobians()
        golden_geometric_jacobian()
    elif os.environ.get("GOLDEN_ONLY") == "voxel":
        golden_voxel_sdf()
    elif os.environ.get("GOLDEN_ONLY") == "points":
        golden_attached_points()
    elif os.environ.get("GOLDEN_ONLY") == "trajectories":
        golden_trajectories()
    elif os.environ.get("GOLDEN_ONLY") == "boxes":
        golden_planning_boxes()
    elif os.environ.get("GOLDEN_ONLY") == "ik500":
        golden_ik_500()
    elif os.environ.get("GOLDEN_ONLY") == "geometric_jacobian":
        golden_geometric_jacobian()
    else:
        golden_fk()
        golden_planning()
        golden_planning_boxes()
        golden_planning_envs()
        golden_stretch_objective()
        golden_ik()
        golden_ik_500()
        golden_trajectories()
        golden_attached_points()
        golden_voxel_sdf()
        golden_link_jacobians()
        golden_geometric_jacobian()
