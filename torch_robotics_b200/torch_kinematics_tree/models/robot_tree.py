"""DifferentiableTree: the reference's kinematic-tree class (models/robot_tree.py:75-492) with the
stateless FK path routed to the sm_100a kernels.

Kept: constructor signature, `compute_forward_kinematics_all_links`, `inverse_kinematics`,
`loss_fn_ik_per_q`, `ik_termination`, the accessors (`get_link_names`, `get_joint_limits`,
`get_joint_limit_array`, `print_link_names`, `update_base_pose`) and the attributes callers read
(`_n_dofs`, `_controlled_joints`, `_name_to_idx_map`, `_bodies`, `link_list`).
Not provided (out of the hot path): the stateful velocity
propagation (`update_kinematic_state`) and the stateful geometric-Jacobian helper built on it.

q must be a float32 CUDA tensor: there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ... import compat, ops
from ..._lib import DeviceModel
from ...descriptor import TreeDescriptor, attach_bodies, attach_points, compile_urdf, quat_wxyz_to_matrix
from ...torch_utils.torch_utils import to_torch, torch_intersect_1d
from ..geometrics.frame import Frame
from ..geometrics.utils import SE3_distance


def convert_link_dict_to_tensor(link_dict, link_list):
    return torch.stack([link_dict[name].get_transform_matrix() for name in link_list], dim=1)


class BodyView:
    """Read-only stand-in for the reference's DifferentiableRigidBody (models/rigid_body.py:66-123):
    the constants of one link and its parent joint."""

    def __init__(self, desc: TreeDescriptor, i: int):
        self.name = desc.link_names[i]
        self.joint_id = i
        self.joint_name = desc.joint_names[i]
        self.joint_type = desc.joint_types[i]
        self.joint_idx = None if desc.q_index[i] < 0 else int(desc.q_index[i])
        self.joint_limits = desc.limits[i]
        self.joint_axis = torch.from_numpy(desc.axis[i:i + 1].copy())
        self.trans = torch.from_numpy(desc.trans[i:i + 1].copy())
        self.fixed_rotation = torch.from_numpy(desc.R_fixed[i:i + 1].copy())
        self.is_root = i == 0
        self.parent_idx = int(desc.parent[i])

    def get_joint_limits(self):
        return self.joint_limits


class DifferentiableTree(torch.nn.Module):

    def __init__(self, model_path: str, name="", link_list=None, device="cpu"):
        super().__init__()
        self.name = name
        self.link_list = link_list
        self._device = device
        self.model_type = str(model_path).split(".")[-1]
        if self.model_type != "urdf":
            raise NotImplementedError(f"{self.model_type} is not supported (the reference's MJCF loader is "
                                      "self-declared broken, models/utils.py:43)")
        self._desc = compile_urdf(model_path, name)
        self._n_point_bodies = 0      # trailing bodies that are attached points (hidden from get_link_names)
        self._bodies = [BodyView(self._desc, i) for i in range(self._desc.n_links)]
        self._n_dofs = self._desc.n_dofs
        self._controlled_joints = list(self._desc.controlled)
        self._name_to_idx_map = self._desc.name_to_idx()
        self._device_models: Dict[int, DeviceModel] = {}
        self._model_generation = 0      # bumped whenever the device model must be rebuilt: dependants key their caches on it

    # ------------------------------------------------------------------------------------ device side
    def device_model(self, device) -> DeviceModel:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"forward kinematics was requested on '{device}': this implementation runs on "
                               "CUDA (sm_100a) only and has no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._device_models:
            self._device_models[idx] = DeviceModel(self._desc, torch.device("cuda", idx), library=self._library_for_current_tree())
        return self._device_models[idx]

    def base_translation(self):
        """The base pose as a pure translation [3] (fp32) - or None when it is the identity or has a rotation."""
        R, t = np.asarray(self._desc.base_R, np.float32), np.asarray(self._desc.base_t, np.float32).reshape(3)
        return t.copy() if np.array_equal(R, np.eye(3, dtype=np.float32)) and np.any(t != 0) else None

    def device_model_in_base_frame(self, device) -> DeviceModel:
        """The tree WITHOUT its base pose (the compiled-in kernels serve it when the robot is a shipped one); callers pull the
        scene back into the base frame instead (descriptor.cost_in_translated_frame)."""
        import copy
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = ("base_frame", idx)
        if key not in self._device_models:
            d = copy.copy(self._desc)
            d.base_R, d.base_t = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
            self._device_models[key] = DeviceModel(d, torch.device("cuda", idx))
        return self._device_models[key]

    def specialize(self, collision_points=None, self_collision_pairs=None, ee_link=None, cache_dir=None, verbose=False):
        """Compile this robot into the unrolled kernel family (specialize.py): afterwards FK, its reverse mode, the link
        Jacobians, the fused IK and - when a collision model is given - the fused planning objective run on kernels
        generated for exactly this tree, as the ten shipped robots do, instead of the table-driven runtime-tree kernels.
        No counterpart in the reference; results follow the same parity bars (tests/test_specialize.py).

        collision_points / self_collision_pairs / ee_link: link names or body indices of the planning objective's
        collision model (RobotBase's link_idxs_for_object_collision_checking, its self-collision pairs and link_name_ee);
        omit all three to specialise the kinematics only.  Takes about a minute of nvcc the first time; the library is
        cached by content (TRK_B200_JIT_DIR, default torch_robotics_b200/csrc/jit).  Attaching links / points or moving
        the base afterwards returns the tree to the default library.  Returns the library's path."""
        from ... import specialize as sp
        to_idx = lambda v: self._name_to_idx_map[v] if isinstance(v, str) else int(v)
        collision = None
        if collision_points is not None or self_collision_pairs is not None or ee_link is not None:
            if ee_link is None:
                raise ValueError("a compile-time collision model needs the end-effector link (ee_link)")
            collision = dict(pt_link=[to_idx(v) for v in (collision_points or [])],
                             pairs=[(to_idx(a), to_idx(b)) for a, b in (self_collision_pairs or [])], ee_body=to_idx(ee_link))
        path = sp.build_specialized_library(self._desc, collision, cache_dir=cache_dir, verbose=verbose)
        self._model_generation += 1          # dependants (PlanningTask's fused cost, the fields' device costs) rebuild on the new library
        self._library = (path, self._model_generation)
        self._device_models.clear()
        return path

    def _library_for_current_tree(self):
        lib_gen = getattr(self, "_library", None)
        return lib_gen[0] if lib_gen is not None and lib_gen[1] == self._model_generation else None

    def _refresh_after_attach(self):
        self._bodies = [BodyView(self._desc, i) for i in range(self._desc.n_links)]
        self._name_to_idx_map = self._desc.name_to_idx()
        self._device_models.clear()
        self._model_generation += 1

    def attach_link(self, name, parent, xyz, rpy=(0.0, 0.0, 0.0)):
        """Append a fixed link `name` under `parent` with the given joint origin - the in-memory equivalent of the
        reference's URDF rewriting for grasped objects (models/robots.py:24-53).  Must precede attach_points()."""
        assert self._n_point_bodies == 0, "attach links before points"
        self._desc, ids = attach_bodies(self._desc, [(name, parent, xyz, rpy)])
        self._refresh_after_attach()
        return ids[0]

    def attach_points(self, points):
        """Rigidly attach points to links: (link name, (x, y, z) in the link frame) pairs become extra fixed bodies
        whose origins are the points' world positions (grasped-object points, multi-sphere link models).  They are
        hidden from get_link_names(); get_all_body_names() lists them after the links.  Returns (names, indices)."""
        self._desc, ids = attach_points(self._desc, points)
        self._n_point_bodies += len(ids)
        self._refresh_after_attach()
        return [self._desc.link_names[i] for i in ids], ids

    def update_base_pose(self, pose_vec):
        """pose_vec: x, y, z, qw, qx, qy, qz (reference: robot_tree.py:133, frame.py:42-51)."""
        p = np.asarray(pose_vec.detach().cpu().numpy() if torch.is_tensor(pose_vec) else pose_vec, np.float32).reshape(-1)
        self._desc.base_t = p[:3].copy()
        self._desc.base_R = quat_wxyz_to_matrix(p[3:7])
        self._device_models.clear()
        self._model_generation += 1

    def reset(self):
        pass

    # ------------------------------------------------------------------------------------ FK
    def compute_forward_kinematics_all_links(self, q: torch.Tensor, return_dict=False, link_list=None):
        """Stateless FK.  q [B, n_dofs] (or [n_dofs]) -> [B, len(link_list), 4, 4], or a dict
        name -> Frame when return_dict=True (reference: robot_tree.py:267-301)."""
        if q.ndim == 1:
            q = q.unsqueeze(0)
        if link_list is None:
            link_list = self.link_list
        if link_list is None:
            link_list = self.get_link_names()
        idx = [self._name_to_idx_map[n] for n in link_list]
        uniq = sorted(set(idx), key=idx.index)
        all_in_order = uniq == list(range(self._desc.n_links))
        home = q.device
        if q.device.type != "cuda" and compat.cpu_tensors_accepted():
            # opt-in compatibility mode (compat.accept_cpu_tensors): the CPU tensor is evaluated on cuda:0 by the same kernels
            q = q.to(torch.device("cuda", torch.cuda.current_device()))
        H = ops.fk(self.device_model(q.device), q, None if all_in_order else uniq)
        if len(uniq) != len(idx):
            H = H[:, [uniq.index(i) for i in idx]]
        if H.device != home:
            H = H.to(home)
        if not return_dict:
            return H
        return {n: Frame(rot=H[:, k, :3, :3], trans=H[:, k, :3, 3], device=home) for k, n in enumerate(link_list)}

    def compute_analytical_jacobian_all_links(self, q: torch.Tensor) -> torch.Tensor:
        """q [B, n_dofs] (or [n_dofs]) -> [B, n_links, 7, n_dofs]: d (position xyz, quaternion wxyz) of every link /
        d q.  The reference gets it from torch.autograd.functional.jacobian over FK + rotation_matrix_to_q
        (robot_tree.py:250-265); here it is one kernel evaluating the closed form (trk_link_jacobian).  Links follow
        self.link_list when set, like the FK call inside the reference's surrogate."""
        if q.ndim == 1:
            q = q.unsqueeze(0)
        J = self.device_model(q.device).link_jacobian(q.detach())
        names = self.link_list if self.link_list is not None else self.get_link_names()
        idx = [self._name_to_idx_map[n] for n in names]
        return J if idx == list(range(self._desc.n_links)) else J[:, idx]

    def compute_forward_kinematics(self, q, qd=None, link_name=None, state_less=True):
        """state_less=True: [B, 1, 4, 4] of `link_name`.  Otherwise (position [B, 3], quaternion wxyz [B, 4]) as the
        reference's stateful call returns them (robot_tree.py:196-216); body velocities are not propagated (qd is unused)."""
        if state_less:
            return self.compute_forward_kinematics_all_links(q, link_list=[link_name])
        H = self.compute_forward_kinematics_all_links(q, link_list=[link_name])[:, 0]
        return H[:, :3, 3], _quaternion_wxyz_of_frame(H[:, :3, :3])

    def compute_forward_kinematics_and_geometric_jacobian(self, q, qd, link_name):
        """(ee_pos [B, 3], ee_rot wxyz [B, 4], lin_jac [B, 3, n_dofs], ang_jac [B, 3, n_dofs]) of `link_name`
        (robot_tree.py:218-248): lin[:, i] = z_i x (p_ee - p_i), ang[:, i] = z_i, one kernel (trk_geometric_jacobian).
        Like the reference, a joint contributes when its index among the URDF's <joint> elements is not larger than that of
        the joint above `link_name`, assuming body k hangs from joint k - 1 (:237-240); z_i is the column of the joint axis'
        non-zero component of the body's world rotation, whatever the axis' sign or the joint type (rigid_body.py:97-99).
        Poses are those of compute_forward_kinematics_all_links.  The reference propagates them statefully here
        (update_kinematic_state), which gives the same poses for positive-axis, limited joints (golden: Panda, UR10, Tiago)
        but rotates negative-axis joints about z without the sign (rigid_body.py:101-106) and clamps continuous joints
        (:218): those two upstream defects are not reproduced."""
        assert q.ndim == 2
        link = self._name_to_idx_map[link_name]
        link_joint = self._desc.urdf_joint_index[link] if link < len(self._desc.urdf_joint_index) else -1
        mask = 0
        for col, body in enumerate(self._controlled_joints):
            if (body - 1) <= link_joint:
                mask |= 1 << col
        pose, lin, ang = self.device_model(q.device).geometric_jacobian(q.detach(), link, mask)
        return pose[:, 9:12], _quaternion_wxyz_of_frame(pose[:, :9].reshape(-1, 3, 3)), lin, ang

    # ------------------------------------------------------------------------------------ IK
    def inverse_kinematics(self, H_target, link_name="ee_link", batch_size=1, max_iters=1000, lr=1e-2, se3_eps=1e-1,
                           q0=None, q0_noise=torch.pi / 8, eps_joint_lim=torch.pi / 100, print_freq=50, debug=False):
        """Batched Adam IK with the reference's loop structure (robot_tree.py:303-384): every seed is
        updated every iteration; stops early when every seed satisfies the termination test."""
        device = H_target.device if torch.is_tensor(H_target) and H_target.is_cuda else torch.device(self._device)
        if H_target.ndim == 2:
            H_target = H_target.unsqueeze(0)
        lower, upper, _, _ = self.get_joint_limit_array()
        lower = to_torch(lower + eps_joint_lim, device=device)
        upper = to_torch(upper - eps_joint_lim, device=device)
        if q0 is None:
            q0 = lower + torch.rand(batch_size, self._n_dofs, device=device) * (upper - lower)
        else:
            q0 = q0 + torch.randn(batch_size, self._n_dofs, device=device) * q0_noise
            q0 = torch.clamp(q0, lower, upper)
            assert q0.shape == (batch_size, self._n_dofs)
        model = self.device_model(device)
        if not debug and max_iters >= 1:      # debug=True keeps the step-by-step torch loop (prints per-iteration state)
            return self._inverse_kinematics_fused(model, q0, H_target, link_name, batch_size, max_iters, lr, se3_eps,
                                                  lower, upper, print_freq)
        q = q0.clone().requires_grad_(True)
        optimizer = torch.optim.Adam([q], lr=lr)
        idx_valid = torch.empty(0, dtype=torch.int64, device=device)
        err_per_q = None
        for i in range(max_iters):
            optimizer.zero_grad()
            idx_valid = self.ik_termination(q, H_target, link_name, lower, upper, se3_eps=se3_eps, debug=debug)
            if idx_valid.nelement() == batch_size:
                if print_freq != -1:
                    print(f"\nIK converged for all joint configurations in {i} iterations")
                break
            err_per_q = self.loss_fn_ik_per_q(q, H_target, link_name, w_se3=1.0, w_joint_limits=300.0,
                                              lower=lower, upper=upper, w_q_rest=1.0, q_rest=None, debug=debug)
            if (i == 0 or (i % print_freq) == 0) and print_freq != -1:
                print(f"\n---> Iter {i}/{max_iters}")
                print(f"Error mean, std: {err_per_q.mean():.3f}, {err_per_q.std():.3f}")
                print(f"idx_valid: {len(idx_valid)}/{batch_size}")
            err_per_q.sum().backward()
            optimizer.step()
        else:
            if print_freq != -1 and err_per_q is not None:
                print("\nIK did not converge for all joint configurations!")
                print(f"idx_valid: {len(idx_valid)}/{batch_size}")
        return q, idx_valid

    def _inverse_kinematics_fused(self, model, q0, H_target, link_name, batch_size, max_iters, lr, se3_eps, lower, upper,
                                  print_freq):
        """The same loop inside one kernel (trk_ik_adam): q and the Adam moments stay in registers; the host
        reads back one per-iteration valid-seed count at the end instead of synchronising every iteration."""
        ee = self._name_to_idx_map[link_name]
        kw = dict(lr=lr, se3_eps=se3_eps, w_se3=1.0, w_joint_limits=300.0)
        want_stats = print_freq != -1
        # The reference stops at the first iteration at which EVERY seed is valid (robot_tree.py:356).  A large batch practically
        # never gets there, so it runs max_iters once; a small batch often converges within tens of iterations, so it is probed
        # with short runs first (the kernel is deterministic: a run of n iterations is the prefix of a longer one) instead of
        # paying max_iters + k iterations for a stop at k.
        probes = [n for n in (64, 256) if n < max_iters] if (batch_size <= 4096 and not want_stats) else []
        for n in probes + [max_iters]:
            r = model.ik_adam(q0, H_target, ee, lower, upper, n_iters=n, want_loss_stats=want_stats, **kw)
            count = r["valid_count"].cpu()
            hit = torch.nonzero(count == batch_size)
            if hit.numel() > 0:
                break
        n_done = max_iters
        if hit.numel() > 0:          # every seed was valid at iteration `it`: the reference stops before updating
            n_done = int(hit[0])
            r2 = model.ik_adam(q0, H_target, ee, lower, upper, n_iters=n_done + 1, n_updates=n_done, **kw)
            r2["loss_sum"], r2["loss_sqsum"] = r["loss_sum"], r["loss_sqsum"]
            r = r2
        if want_stats:
            ls, lq = r["loss_sum"].double().cpu(), r["loss_sqsum"].double().cpu()
            for i in range(min(n_done, max_iters)):
                if i == 0 or (i % print_freq) == 0:
                    mean = ls[i] / batch_size
                    std = ((lq[i] - batch_size * mean * mean).clamp_min(0) / max(batch_size - 1, 1)).sqrt()
                    print(f"\n---> Iter {i}/{max_iters}")
                    print(f"Error mean, std: {mean:.3f}, {std:.3f}")
                    print(f"idx_valid: {int(count[i])}/{batch_size}")
            if hit.numel() > 0:
                print(f"\nIK converged for all joint configurations in {n_done} iterations")
            else:
                print("\nIK did not converge for all joint configurations!")
                print(f"idx_valid: {int(count[max_iters - 1])}/{batch_size}")
        idx_valid = torch.nonzero(r["valid"]).reshape(-1)
        return r["q"].requires_grad_(True), idx_valid

    def loss_fn_ik_per_q(self, q, H_target, link_name, w_se3=1.0, w_joint_limits=1.0, lower=None, upper=None,
                         w_q_rest=1.0, q_rest=None, debug=False):
        H = self.compute_forward_kinematics_all_links(q, link_list=[link_name]).squeeze(1)
        err_se3 = SE3_distance(H, H_target, w_pos=1.0, w_rot=1.0)
        below = (q < lower).to(q.dtype)
        above = (q > upper).to(q.dtype)
        err_limits = ((lower - q).pow(2) * below).sum(-1) + ((upper - q).pow(2) * above).sum(-1)
        err_rest = 0.0 if q_rest is None else torch.linalg.norm(q - q_rest, dim=-1)
        return w_se3 * err_se3 + w_joint_limits * err_limits + w_q_rest * err_rest

    def ik_termination(self, q, H_target, link_name, lower, upper, se3_eps=1e-1, debug=False):
        in_limits = torch.atleast_1d(torch.argwhere(torch.all((q >= lower) & (q <= upper), dim=-1)).squeeze())
        with torch.no_grad():
            H = self.compute_forward_kinematics_all_links(q.detach(), link_list=[link_name]).squeeze(1)
            err_se3 = SE3_distance(H, H_target, w_pos=1.0, w_rot=1.0)
        close = torch.atleast_1d(torch.argwhere(err_se3 < se3_eps).squeeze())
        return torch_intersect_1d(in_limits, close)

    # ------------------------------------------------------------------------------------ accessors
    def get_joint_limits(self) -> List[Optional[Dict[str, float]]]:
        return [self._bodies[i].get_joint_limits() for i in self._controlled_joints]

    def get_joint_limit_array(self) -> Tuple[np.ndarray, ...]:
        lims = self.get_joint_limits()
        lo = np.array([d["lower"] for d in lims])
        hi = np.array([d["upper"] for d in lims])
        vel = np.array([d["velocity"] for d in lims])
        return lo, hi, -vel, vel

    def get_link_names(self) -> List[str]:
        return list(self._desc.link_names[:self._desc.n_links - self._n_point_bodies])

    def get_all_body_names(self) -> List[str]:
        """Links followed by the attached points."""
        return list(self._desc.link_names)

    def print_link_names(self) -> None:
        for n in self.get_link_names():
            print(n)


def _quaternion_wxyz_of_frame(R):
    """Frame.get_quaternion (geometrics/frame.py:87-113, an xyzw quaternion from the homogeneous matrix, un-normalised branch
    values scaled by 0.5 / sqrt(t)) followed by q_convert_wxyz (quaternion.py:240-242), vectorised over the batch."""
    m = lambda i, j: R[:, i, j]
    tr = m(0, 0) + m(1, 1) + m(2, 2) + 1.0              # trace of the 4 x 4 homogeneous matrix
    # branch 0: trace > 1
    q0 = torch.stack([m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1), tr], dim=-1)
    t0 = tr
    # otherwise the largest diagonal element picks (i, j, k)
    d = torch.stack([m(0, 0), m(1, 1), m(2, 2)], dim=-1)
    i = torch.zeros_like(tr, dtype=torch.long)
    i = torch.where(d[:, 1] > d[:, 0], torch.ones_like(i), i)
    i = torch.where(d[:, 2] > d.gather(1, i[:, None])[:, 0], torch.full_like(i, 2), i)
    j, k = (i + 1) % 3, (i + 2) % 3
    ar = torch.arange(R.shape[0], device=R.device)
    e = lambda a, b: R[ar, a, b]
    tn = e(i, i) - (e(j, j) + e(k, k)) + 1.0
    q1 = torch.zeros_like(q0)
    q1[ar, i] = tn
    q1[ar, j] = e(i, j) + e(j, i)
    q1[ar, k] = e(k, i) + e(i, k)
    q1[:, 3] = e(k, j) - e(j, k)
    first = tr > 1.0
    t = torch.where(first, t0, tn)
    q = torch.where(first[:, None], q0, q1) * (0.5 / torch.sqrt(t))[:, None]
    return torch.stack([q[:, 3], q[:, 0], q[:, 1], q[:, 2]], dim=-1)
