"""Host-side model compiler: URDF text -> flat kinematic-tree descriptor arrays.

This is the once-per-model step that the reference spreads over
`URDFRobotModel.get_body_parameters` (torch_kinematics_tree/models/utils.py:199-313),
`DifferentiableRigidBody.__init__` (models/rigid_body.py:74-123) and
`DifferentiableTree.__init__` (models/robot_tree.py:77-126).  The output is what the CUDA
kernels consume through `trk_model_create` (include/trk_b200.h): one row per link, in URDF
`<link>` order ("body order"), plus a parents-first processing order.

Reference semantics that are reproduced on purpose (SURVEY.md section 8a "quirks"):
  * body 0 (= first <link>) is the root: zero offset, fixed (models/utils.py:204-211);
  * q column i drives the i-th non-fixed body in <link> order (robot_tree.py:111-115);
  * R_fixed = (Rz(yaw) @ Ry(pitch)) @ Rx(roll) evaluated with fp32 torch ops (rigid_body.py:93),
    so e.g. cos(fp32(-1.57079632679)) = -4.37e-8 is kept, not rounded to 0;
  * revolute axis: |a_x|==1 -> x, elif |a_y|==1 -> y, else z with sign(a_z)
    (rigid_body.py:163-168) - an oblique axis turns about z, a zero axis does not turn at all;
  * prismatic translation uses the raw axis vector, not rotated by the joint rpy
    (rigid_body.py:176-181);
  * FK clamps q to [lower, upper] unless the joint is 'continuous' or has no <limit>
    (rigid_body.py:157-160); continuous joints with a <limit> store +-pi (models/utils.py:240-242).
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

KIND_FIXED, KIND_REVOLUTE, KIND_PRISMATIC = 0, 1, 2
MAX_LINKS = 64   # must match TRK_MAX_LINKS in include/trk_b200.h
MAX_DOFS = 32    # must match TRK_MAX_DOFS


def _floats(text: Optional[str]) -> List[float]:
    if text is None:
        return [0.0, 0.0, 0.0]
    return [float(v) for v in text.split()]


def _axis_rot(k: int, angle: torch.Tensor) -> torch.Tensor:
    """fp32 [1,3,3] axis rotation built the way spatial_vector.py:8-47 does (zeros + cos/sin)."""
    c, s = torch.cos(angle), torch.sin(angle)
    R = torch.zeros((1, 3, 3), dtype=torch.float32)
    if k == 0:
        R[:, 0, 0] = 1.0
        R[:, 1, 1] = c
        R[:, 1, 2] = -s
        R[:, 2, 1] = s
        R[:, 2, 2] = c
    elif k == 1:
        R[:, 0, 0] = c
        R[:, 0, 2] = s
        R[:, 1, 1] = 1.0
        R[:, 2, 0] = -s
        R[:, 2, 2] = c
    else:
        R[:, 0, 0] = c
        R[:, 0, 1] = -s
        R[:, 1, 0] = s
        R[:, 1, 1] = c
        R[:, 2, 2] = 1.0
    return R


def fixed_rotation_from_rpy(rpy) -> np.ndarray:
    """(Rz(yaw) @ Ry(pitch)) @ Rx(roll) in fp32 torch CPU arithmetic (rigid_body.py:88-93)."""
    r = torch.tensor(rpy, dtype=torch.float32)
    R = (_axis_rot(2, r[2]) @ _axis_rot(1, r[1])) @ _axis_rot(0, r[0])
    return R[0].numpy().copy()


@dataclass
class TreeDescriptor:
    name: str
    link_names: List[str]
    joint_names: List[str]
    joint_types: List[str]
    parent: np.ndarray      # int32 [L], -1 for the root
    kind: np.ndarray        # int32 [L]  KIND_*
    axis_k: np.ndarray      # int32 [L]  0/1/2 rotation axis of revolute joints
    axis_s: np.ndarray      # float32 [L] sign applied to q (-1, 0, +1)
    axis: np.ndarray        # float32 [L,3] raw <axis xyz>
    q_index: np.ndarray     # int32 [L]  column of q, -1 for fixed bodies
    R_fixed: np.ndarray     # float32 [L,3,3]
    trans: np.ndarray       # float32 [L,3]
    lo: np.ndarray          # float32 [L]  clamp bounds used by FK
    hi: np.ndarray          # float32 [L]
    has_lim: np.ndarray     # int32 [L]  1 when FK clamps this joint
    topo: np.ndarray        # int32 [L]  processing order, parents before children
    depth: np.ndarray       # int32 [L]
    limits: List[Optional[Dict[str, float]]]   # per body, as models/utils.py:234-243 stores them
    controlled: List[int] = field(default_factory=list)   # body index per q column
    urdf_joint_index: List[int] = field(default_factory=list)   # per body: position of its parent joint among the <joint> elements, -1 for the root
    base_R: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=np.float32))
    base_t: np.ndarray = field(default_factory=lambda: np.zeros(3, dtype=np.float32))

    @property
    def n_links(self) -> int:
        return len(self.link_names)

    @property
    def n_dofs(self) -> int:
        return len(self.controlled)

    def name_to_idx(self) -> Dict[str, int]:
        # later duplicates win, like the dict assignment at robot_tree.py:119
        return {n: i for i, n in enumerate(self.link_names)}

    def signature(self) -> bytes:
        """Byte string identifying the kinematics (used to pick a compiled-in specialisation)."""
        parts = [self.parent, self.kind, self.axis_k, self.axis_s, self.axis, self.q_index,
                 self.R_fixed, self.trans, self.lo, self.hi, self.has_lim]
        return b"".join(np.ascontiguousarray(p).tobytes() for p in parts)


def compile_urdf(path: str, name: str = "") -> TreeDescriptor:
    root = ET.parse(str(path)).getroot()
    link_names = [el.get("name") for el in root if el.tag == "link"]
    joints = [el for el in root if el.tag == "joint"]
    L = len(link_names)
    if L == 0:
        raise ValueError(f"{path}: no <link> elements")
    if L > MAX_LINKS:
        raise ValueError(f"{path}: {L} links exceeds the compiled maximum of {MAX_LINKS}")

    def joint_of(child_name):   # first joint whose child is this link (models/utils.py:188-192)
        for j in joints:
            if j.find("child").get("link") == child_name:
                return j
        return None

    name_to_idx = {n: i for i, n in enumerate(link_names)}
    parent = np.full(L, -1, np.int32)
    kind = np.zeros(L, np.int32)
    axis_k = np.full(L, 2, np.int32)
    axis_s = np.zeros(L, np.float32)
    axis = np.zeros((L, 3), np.float32)
    q_index = np.full(L, -1, np.int32)
    R_fixed = np.tile(np.eye(3, dtype=np.float32), (L, 1, 1))
    trans = np.zeros((L, 3), np.float32)
    lo = np.zeros(L, np.float32)
    hi = np.zeros(L, np.float32)
    has_lim = np.zeros(L, np.int32)
    limits: List[Optional[Dict[str, float]]] = [None] * L
    joint_names = ["base_joint"] * L
    joint_types = ["fixed"] * L
    controlled: List[int] = []
    urdf_joint_index = [-1] * L

    if joint_of(link_names[0]) is not None:
        raise ValueError(f"{path}: the first <link> ({link_names[0]}) must be the tree root "
                         "(the reference roots the recursion there, robot_tree.py:293)")

    for i in range(1, L):
        j = joint_of(link_names[i])
        if j is None:
            raise ValueError(f"{path}: link {link_names[i]} has no parent joint (multiple roots)")
        jt = j.get("type")
        urdf_joint_index[i] = joints.index(j)
        joint_names[i] = j.get("name")
        joint_types[i] = jt
        parent[i] = name_to_idx[j.find("parent").get("link")]
        org = j.find("origin")
        xyz = _floats(org.get("xyz")) if org is not None else [0.0, 0.0, 0.0]
        rpy = _floats(org.get("rpy")) if org is not None else [0.0, 0.0, 0.0]
        trans[i] = np.asarray(xyz, np.float32)
        R_fixed[i] = fixed_rotation_from_rpy(rpy)
        ax = j.find("axis")
        if ax is not None:
            axis[i] = np.asarray(_floats(ax.get("xyz")), np.float32)
        if jt == "fixed":
            continue
        if jt not in ("revolute", "continuous", "prismatic"):
            raise NotImplementedError(f"{path}: joint type '{jt}' (rigid_body.py:183 raises too)")
        q_index[i] = len(controlled)
        controlled.append(i)
        if len(controlled) > MAX_DOFS:
            raise ValueError(f"{path}: more than {MAX_DOFS} degrees of freedom")
        kind[i] = KIND_PRISMATIC if jt == "prismatic" else KIND_REVOLUTE
        a = axis[i]
        if abs(a[0]) == 1:
            axis_k[i], axis_s[i] = 0, np.sign(a[0])
        elif abs(a[1]) == 1:
            axis_k[i], axis_s[i] = 1, np.sign(a[1])
        else:
            axis_k[i], axis_s[i] = 2, np.sign(a[2])
        lim = j.find("limit")
        if lim is not None:
            d = {"effort": float(lim.get("effort", 0.0)), "lower": float(lim.get("lower", 0.0)),
                 "upper": float(lim.get("upper", 0.0)), "velocity": float(lim.get("velocity", 0.0))}
            if jt == "continuous":
                d["lower"], d["upper"] = -math.pi, math.pi
            limits[i] = d
            if jt != "continuous":
                has_lim[i] = 1
                lo[i], hi[i] = np.float32(d["lower"]), np.float32(d["upper"])

    # parents-first processing order (stable: body order among ready links) and depth
    depth = np.full(L, -1, np.int32)
    depth[0] = 0
    topo = [0]
    placed = {0}
    while len(topo) < L:
        progressed = False
        for i in range(1, L):
            if i not in placed and int(parent[i]) in placed:
                placed.add(i)
                topo.append(i)
                depth[i] = depth[parent[i]] + 1
                progressed = True
        if not progressed:
            missing = [link_names[i] for i in range(L) if i not in placed]
            raise ValueError(f"{path}: links not reachable from the root: {missing}")

    return TreeDescriptor(
        name=name or root.get("name", ""), link_names=link_names, joint_names=joint_names,
        joint_types=joint_types, parent=parent, kind=kind, axis_k=axis_k, axis_s=axis_s, axis=axis,
        q_index=q_index, R_fixed=R_fixed, trans=trans, lo=lo, hi=hi, has_lim=has_lim,
        topo=np.asarray(topo, np.int32), depth=depth, limits=limits, controlled=controlled, urdf_joint_index=urdf_joint_index)


def attach_bodies(desc: TreeDescriptor, bodies) -> "tuple[TreeDescriptor, list]":
    """Descriptor with extra FIXED bodies appended after the existing ones.

    bodies: iterable of (name, parent link name or body index, xyz, rpy-or-None).  A body with rpy=None and offset
    xyz has its origin at R_w[parent] @ xyz + p_w[parent] - exactly what Frame.transform_point computes for
    grasped-object points (robots/robot_panda.py:155-159, geometrics/frame.py:116) and what a per-link multi-sphere
    collision model (data/configs/*_sphere_config.yaml) needs.  A body with an rpy is what the reference gets by
    rewriting the URDF with a fixed joint (models/robots.py:24-53).  Because they are ordinary bodies, every kernel
    (FK, fused objective, reverse mode) handles them without special code: a force on an attached point produces
    the torque x_point x f on the chain automatically.  Parents may be bodies added earlier in the same call.
    Returns (new descriptor, body indices of the additions)."""
    import copy
    d = copy.deepcopy(desc)
    new_ids = []
    for name, link, xyz, rpy in bodies:
        idx = d.name_to_idx()
        n0 = d.n_links
        if n0 + 1 > MAX_LINKS:
            raise ValueError(f"more than the compiled maximum of {MAX_LINKS} bodies after attaching {name!r}")
        b = idx.get(link, -1) if isinstance(link, str) else int(link)
        if not 0 <= b < n0:
            raise ValueError(f"attached body {name!r}: unknown parent link {link!r}")
        R = np.eye(3, dtype=np.float32) if rpy is None else fixed_rotation_from_rpy([float(v) for v in rpy])
        new_ids.append(n0)
        d.link_names.append(name)
        d.joint_names.append(name + "_fixed_joint")
        d.joint_types.append("fixed")
        d.limits.append(None)
        d.parent = np.append(d.parent, b).astype(np.int32)
        d.kind = np.append(d.kind, KIND_FIXED).astype(np.int32)
        d.axis_k = np.append(d.axis_k, 2).astype(np.int32)
        d.axis_s = np.append(d.axis_s, 0).astype(np.float32)
        d.axis = np.concatenate([d.axis, np.zeros((1, 3), np.float32)]).astype(np.float32)
        d.q_index = np.append(d.q_index, -1).astype(np.int32)
        d.R_fixed = np.concatenate([d.R_fixed, R[None].astype(np.float32)]).astype(np.float32)
        d.trans = np.concatenate([d.trans, np.asarray(xyz, np.float32).reshape(1, 3)]).astype(np.float32)
        d.lo = np.append(d.lo, 0).astype(np.float32)
        d.hi = np.append(d.hi, 0).astype(np.float32)
        d.has_lim = np.append(d.has_lim, 0).astype(np.int32)
        d.topo = np.append(d.topo, n0).astype(np.int32)
        d.depth = np.append(d.depth, d.depth[b] + 1).astype(np.int32)
    return d, new_ids


def attach_points(desc: TreeDescriptor, points, suffix: str = "__pt") -> "tuple[TreeDescriptor, list]":
    """attach_bodies() for plain points: iterable of (link name or body index, (x, y, z) in the link frame)."""
    names = desc.link_names
    idx = desc.name_to_idx()
    bodies = []
    for k, (link, off) in enumerate(points):
        base = link if isinstance(link, str) else names[int(link)]
        if isinstance(link, str) and link not in idx:
            raise ValueError(f"attached point {k}: unknown link {link!r}")
        bodies.append((f"{base}{suffix}{k}", link, off, None))
    return attach_bodies(desc, bodies)


def point_descriptor() -> TreeDescriptor:
    """One fixed body at the world origin: the 'model' used when the cost-stage kernel is run on free points (SDF grid
    precomputation and standalone grid queries, environments/grid_map_sdf.py)."""
    z = lambda dt: np.zeros(1, dt)
    return TreeDescriptor(
        name="point", link_names=["world"], joint_names=["world_joint"], joint_types=["fixed"],
        parent=np.array([-1], np.int32), kind=np.array([KIND_FIXED], np.int32), axis_k=np.array([2], np.int32), axis_s=z(np.float32),
        axis=np.zeros((1, 3), np.float32), q_index=np.array([-1], np.int32), R_fixed=np.eye(3, dtype=np.float32)[None].copy(),
        trans=np.zeros((1, 3), np.float32), lo=z(np.float32), hi=z(np.float32), has_lim=z(np.int32), topo=z(np.int32),
        depth=z(np.int32), limits=[None], controlled=[])


def load_sphere_config(path) -> "list[tuple[str, list, float]]":
    """[(link name, [x, y, z], radius), ...] from a data/configs/**/*_sphere_config.json file (converted from the
    reference's YAML by tools/convert_sphere_config.py)."""
    import json
    with open(path) as f:
        return [(s["link"], s["offset"], float(s["radius"])) for s in json.load(f)["spheres"]]


# ----------------------------------------------------------------------------------------------
# Planning-objective descriptor (collision points, obstacle spheres, self pairs, ws bounds, EE)
# ----------------------------------------------------------------------------------------------
MAX_POINTS = 64     # TRK_MAX_POINTS
MAX_SPHERES = 256   # TRK_MAX_SPHERES
MAX_OBJECTS = 16    # TRK_MAX_OBJECTS
MAX_PAIRS = 128     # TRK_MAX_PAIRS
MAX_BOXES = 64      # TRK_MAX_BOXES


@dataclass
class CostDescriptor:
    """Flat description of the objective evaluated on the FK result.

    Mirrors what PlanningTask wires together (tasks/tasks.py:44-86) from RobotPanda's collision
    model (robots/robot_panda.py:47-127, robots/robot_base.py:57-141) and the environment's
    object list (environments/env_base.py:75-88):

      self term  : sum_p ( margin_p - |x_i - x_j| )                distance_fields.py:194-208,113-124
      object term: sum_s max_o ( (r_s + cutoff) - min_k(|x_s - c_k| - r_k) )   :307-316, primitives.py:108-112
      ws term    : sum_s max_6 ( (r_s + cutoff) - halfspace distance )          :326-332
      ee term    : (w_rot (1 - (tr(R R_t^T) - 1)/2) + w_pos |p - p_t|)^(2 if square)   :347-356, utils.py:130-154

    All margins are fp32 values computed with the same fp32 additions the reference performs.
    """
    # collision points (object + ws terms): one point at the origin of `pt_link[s]`
    pt_link: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    pt_margin_obj: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))   # r_s + cutoff (object term)
    pt_margin_ws: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))    # r_s + cutoff (ws term)
    # obstacle objects: object o owns spheres [obj_start[o], obj_start[o+1])
    obj_start: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    obj_R: np.ndarray = field(default_factory=lambda: np.zeros((0, 3, 3), np.float32))   # object rotation (world <- object)
    obj_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    obj_identity: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    sph_center: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    sph_radius: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    # boxes: object o also owns boxes [obj_box_start[o], obj_box_start[o+1]) (primitives.py:198-334)
    obj_box_start: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    box_center: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    box_half: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    box_radius: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))     # MultiBoxField rounding radius
    box_rounded: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))      # 1 rounded, 0 sharp
    # precomputed voxel SDF (grid_map_sdf.py): at most one object is a grid; see VoxelSDF
    grid_obj: int = -1
    grid_n: np.ndarray = field(default_factory=lambda: np.zeros(3, np.int32))
    grid_lo: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    grid_extent: np.ndarray = field(default_factory=lambda: np.ones(3, np.float32))
    grid_data: object = None      # [n0, n1, n2, 4] fp32 (sdf, grad): torch tensor (any device) or numpy array
    # workspace box
    ws_enabled: int = 0
    ws_min: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    ws_max: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    # self-collision pairs (body indices) and their margins
    pair_i: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    pair_j: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    pair_margin: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    # end-effector SE(3) term
    ee_enabled: int = 0
    ee_link: int = 0
    ee_w_pos: float = 1.0
    ee_w_rot: float = 1.0
    ee_square: int = 1
    ee_target: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    # relu on (margin - sdf) per term: self, object, ws  (distance_fields.py:117)
    clamp_self: int = 0
    clamp_obj: int = 0
    clamp_ws: int = 0
    # weights of the four terms in the total: self, object, ws, ee
    weights: np.ndarray = field(default_factory=lambda: np.ones(4, np.float32))

    @property
    def n_points(self) -> int:
        return int(self.pt_link.shape[0])

    @property
    def n_objects(self) -> int:
        return int(self.obj_start.shape[0]) - 1

    @property
    def n_spheres(self) -> int:
        return int(self.sph_radius.shape[0])

    @property
    def n_pairs(self) -> int:
        return int(self.pair_i.shape[0])

    @property
    def n_boxes(self) -> int:
        return int(self.box_radius.shape[0])

    def validate(self, n_links: int) -> None:
        if self.n_points > MAX_POINTS or self.n_spheres > MAX_SPHERES or self.n_objects > MAX_OBJECTS \
                or self.n_pairs > MAX_PAIRS or self.n_boxes > MAX_BOXES:
            raise ValueError("cost descriptor exceeds compiled table sizes")
        for arr in (self.pt_link, self.pair_i, self.pair_j):
            if arr.size and (arr.min() < 0 or arr.max() >= n_links):
                raise ValueError("cost descriptor references a link outside the model")
        if self.ee_enabled and not (0 <= self.ee_link < n_links):
            raise ValueError("ee_link outside the model")


# ----------------------------------------------------------------------------------------------
# builders
# ----------------------------------------------------------------------------------------------
def quat_wxyz_to_matrix(q) -> np.ndarray:
    """fp32 rotation matrix of a (not necessarily unit) wxyz quaternion, same arithmetic as
    geometrics/quaternion.py:102-120."""
    q = torch.as_tensor(np.asarray(q, np.float32)).reshape(4)
    w, x, y, z = torch.unbind(q, -1)
    dc = 2.0 / (q ** 2).sum(-1)
    o = torch.stack((1 - dc * (y * y + z * z), dc * (x * y - z * w), dc * (x * z + y * w),
                     dc * (x * y + z * w), 1 - dc * (x * x + z * z), dc * (y * z - x * w),
                     dc * (x * z - y * w), dc * (y * z + x * w), 1 - dc * (x * x + y * y)), dim=-1)
    return o.reshape(3, 3).numpy().copy()


def _f32(x) -> np.ndarray:
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, np.float32))


@dataclass
class VoxelSDF:
    """A precomputed signed-distance grid standing in for an object list (GridMapSDF, grid_map_sdf.py:9-114).
    data: [n0, n1, n2, 4] fp32 = (sdf, d/dx, d/dy, d/dz) at the cell's sample point; lo/extent: limits[0] and
    |limits[1] - limits[0]|.  Pass an instance in build_cost_descriptor(objects=[...]) in place of a primitive tuple."""
    n: tuple
    lo: np.ndarray
    extent: np.ndarray
    data: object

    def host_data(self) -> np.ndarray:
        d = self.data
        return np.ascontiguousarray(d.detach().cpu().numpy() if torch.is_tensor(d) else d, np.float32)


def grid_cell_index(x, lo, extent, n):
    """Cell of the points x [..., 3] exactly as GridMapSDF.get_sdf computes it (grid_map_sdf.py:96-101), fp32 torch."""
    x = torch.as_tensor(x, dtype=torch.float32)
    nf = torch.as_tensor(np.asarray(n, np.int64))
    idx = ((x - torch.as_tensor(lo, dtype=torch.float32)) / torch.as_tensor(extent, dtype=torch.float32) * nf).floor().to(torch.int)
    hi = (nf - 1).to(torch.int)
    return idx.clamp(torch.zeros_like(hi), hi)


def cost_in_translated_frame(cost: "CostDescriptor", t) -> "CostDescriptor":
    """The same objective seen from a frame translated by `t` (frame origin at world point t, axes parallel to the world's):
    every world-fixed quantity is shifted by -t.  Distances, hence costs, gradients w.r.t. q, occupancy and arg-min, are those
    of the world-frame evaluation up to one fp32 rounding of the shifted constants.  Used for a robot whose base pose is a pure
    translation (DifferentiableTree.update_base_pose, reference robot_tree.py:133): the tree keeps its own base frame - and with
    it the unrolled kernels of the compiled-in family - while the scene is pulled back (tasks/tasks.py:_fused_cost).
    A rotated base cannot be served this way: the workspace box would no longer be axis-aligned."""
    import copy
    t = np.asarray(t, np.float32).reshape(3)
    c = copy.copy(cost)
    n_obj = len(c.obj_identity)
    c.sph_center, c.box_center, c.obj_pos = c.sph_center.copy(), c.box_center.copy(), c.obj_pos.copy()
    for o in range(n_obj):
        if c.obj_identity[o]:      # world-frame primitives: shift them (the object keeps the identity pose and its fast paths)
            c.sph_center[c.obj_start[o]:c.obj_start[o + 1]] -= t
            c.box_center[c.obj_box_start[o]:c.obj_box_start[o + 1]] -= t
            if o == c.grid_obj:
                c.grid_lo = (np.asarray(c.grid_lo, np.float32) - t).astype(np.float32)
        else:                      # posed object: x' = (x - pos) @ R, so only its position moves
            c.obj_pos[o] -= t
    if c.ws_enabled:
        c.ws_min, c.ws_max = (np.asarray(c.ws_min, np.float32) - t).astype(np.float32), (np.asarray(c.ws_max, np.float32) - t).astype(np.float32)
    if getattr(c, "ee_enabled", 0):
        tgt = np.asarray(c.ee_target, np.float32).reshape(4, 4).copy()
        tgt[:3, 3] -= t
        c.ee_target = tgt
    return c


def build_cost_descriptor(
        pt_link=(), pt_radius=(), cutoff_obj=0.0, cutoff_ws=None,
        objects=(), ws_min=None, ws_max=None,
        self_pairs=(), self_margin=(),
        ee_link=None, ee_target=None, ee_w_pos=1.0, ee_w_rot=1.0, ee_square=True,
        clamp_self=False, clamp_obj=False, clamp_ws=False, weights=(1.0, 1.0, 1.0, 1.0)) -> CostDescriptor:
    """Assemble a CostDescriptor.

    pt_link/pt_radius: collision points (body index, sphere radius) - RobotBase's
        link_idxs_for_object_collision_checking / link_margins_..._tensor (robot_base.py:57-85).
    cutoff_*: the fields' cutoff_margin; the per-point margin is the fp32 sum radius + cutoff
        (distance_fields.py:113).
    objects: iterable of (centers [K,3], radii [K], pos [3] | None, ori_wxyz [4] | None[, boxes]) - one entry
        per ObjectField; several primitive sphere fields of one object are simply concatenated
        (min over fields of min over spheres, primitives.py:402-405).  `boxes` is an optional list of
        (centers [KB,3], sizes [KB,3], rounded: bool) for MultiBoxField (rounded) / MultiSharpBoxField entries.
    self_pairs: (body_i, body_j) pairs, self_margin their margins (robot_base.py:105-141).
    """
    c = CostDescriptor()
    c.pt_link = np.ascontiguousarray(np.asarray(pt_link, np.int32).reshape(-1))
    rad = torch.as_tensor(_f32(pt_radius).reshape(-1))
    cutoff_ws = cutoff_obj if cutoff_ws is None else cutoff_ws
    c.pt_margin_obj = (rad + cutoff_obj).numpy().astype(np.float32) if rad.numel() else np.zeros(0, np.float32)
    c.pt_margin_ws = (rad + cutoff_ws).numpy().astype(np.float32) if rad.numel() else np.zeros(0, np.float32)
    starts, Rs, poss, idents, cs, rs = [0], [], [], [], [], []
    bstarts, bc, bh, br, bf = [0], [], [], [], []
    for oi, obj in enumerate(objects):
        if isinstance(obj, VoxelSDF):
            if c.grid_obj >= 0:
                raise ValueError("at most one voxel SDF per objective (EnvBase builds one for its fixed objects)")
            c.grid_obj, c.grid_n = oi, np.asarray(obj.n, np.int32).reshape(3)
            c.grid_lo, c.grid_extent, c.grid_data = _f32(obj.lo).reshape(3), _f32(obj.extent).reshape(3), obj.data
            obj = (np.zeros((0, 3), np.float32), np.zeros(0, np.float32), None, None)
        centers, radii = _f32(obj[0]).reshape(-1, 3), _f32(obj[1]).reshape(-1)
        nb = 0
        for bcen, bsize, rounded in (obj[4] if len(obj) > 4 and obj[4] is not None else []):
            sizes = torch.as_tensor(_f32(bsize).reshape(-1, 3))
            bc.append(_f32(bcen).reshape(-1, 3))
            bh.append((sizes / 2).numpy())                                        # primitives.py:214
            br.append((torch.min(sizes, dim=-1)[0] * 0.15).numpy() if rounded else np.zeros(len(sizes), np.float32))   # :325
            bf.append(np.full(len(sizes), 1 if rounded else 0, np.int32))
            nb += len(sizes)
        bstarts.append(bstarts[-1] + nb)
        pos = None if len(obj) < 3 or obj[2] is None else _f32(obj[2]).reshape(3)
        ori = None if len(obj) < 4 or obj[3] is None else _f32(obj[3]).reshape(4)
        ident = (pos is None or not np.any(pos)) and (ori is None or (ori[0] == 1 and not np.any(ori[1:])))
        Rs.append(np.eye(3, dtype=np.float32) if ori is None else quat_wxyz_to_matrix(ori))
        poss.append(np.zeros(3, np.float32) if pos is None else pos)
        idents.append(1 if ident else 0)
        cs.append(centers)
        rs.append(radii)
        starts.append(starts[-1] + len(radii))
    c.obj_start = np.asarray(starts, np.int32)
    c.obj_R = np.stack(Rs).astype(np.float32) if Rs else np.zeros((0, 3, 3), np.float32)
    c.obj_pos = np.stack(poss).astype(np.float32) if poss else np.zeros((0, 3), np.float32)
    c.obj_identity = np.asarray(idents, np.int32)
    c.sph_center = np.concatenate(cs).astype(np.float32) if cs else np.zeros((0, 3), np.float32)
    c.sph_radius = np.concatenate(rs).astype(np.float32) if rs else np.zeros(0, np.float32)
    c.obj_box_start = np.asarray(bstarts, np.int32)
    if bc:
        c.box_center, c.box_half = np.concatenate(bc).astype(np.float32), np.concatenate(bh).astype(np.float32)
        c.box_radius, c.box_rounded = np.concatenate(br).astype(np.float32), np.concatenate(bf).astype(np.int32)
    if ws_min is not None:
        c.ws_enabled, c.ws_min, c.ws_max = 1, _f32(ws_min).reshape(3), _f32(ws_max).reshape(3)
    pairs = np.asarray(self_pairs, np.int32).reshape(-1, 2)
    c.pair_i, c.pair_j = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
    m = _f32(self_margin).reshape(-1)
    c.pair_margin = np.broadcast_to(m, (len(pairs),)).astype(np.float32).copy() if len(pairs) else np.zeros(0, np.float32)
    if ee_link is not None:
        c.ee_enabled, c.ee_link = 1, int(ee_link)
        c.ee_w_pos, c.ee_w_rot, c.ee_square = float(ee_w_pos), float(ee_w_rot), int(bool(ee_square))
        if ee_target is not None:
            c.ee_target = _f32(ee_target).reshape(4, 4)
    c.clamp_self, c.clamp_obj, c.clamp_ws = int(clamp_self), int(clamp_obj), int(clamp_ws)
    c.weights = _f32(weights).reshape(4)
    return c
