import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "torch_robotics_b200", "data", "urdf", "robots")

URDF_FILES = dict(
    panda="franka_description/robots/panda_arm_no_gripper.urdf",
    panda_gripper="franka_description/robots/panda_arm_hand.urdf",
    ur10="ur10/urdf/ur10.urdf",
    kuka_iiwa="kuka_iiwa/urdf/iiwa7.urdf",
    stretch="habitat_stretch/urdf/hab_stretch.urdf",
    tiago_dual_holo="tiago_dual_description/tiago_dual_holobase_minimal.urdf",
    tiago_dual_holo_move="tiago_dual_description/tiago_dual_holobase_minimal_holonomic.urdf",
    shadow_hand="shadow_hand/shadow_hand.urdf",
    allegro_hand="allegro_hand/allegro_hand.urdf",
    planar_2link="planar_manipulators/urdf/2_link_planar.urdf",
)
ROBOT_NAMES = list(URDF_FILES)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


_DESC_CACHE = {}


def get_desc(robot):
    from torch_robotics_b200.descriptor import compile_urdf
    if robot not in _DESC_CACHE:
        _DESC_CACHE[robot] = compile_urdf(os.path.join(DATA, URDF_FILES[robot]), name=robot)
    return _DESC_CACHE[robot]


def rel_err(a, b):
    """Norm-wise relative error per tensor: |a-b|_inf / |b|_inf (tolerance definition)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def planning_cost_from_golden(g, with_ee=True):
    """CostDescriptor for a tests/golden/planning_*.npz case (Panda collision model)."""
    from torch_robotics_b200.descriptor import build_cost_descriptor
    objects = []
    for o in range(int(g["n_obj"])):
        boxes = []
        if f"obj{o}_n_box_fields" in g.files:
            for b in range(int(g[f"obj{o}_n_box_fields"])):
                boxes.append((g[f"obj{o}_box{b}_centers"], g[f"obj{o}_box{b}_sizes"], bool(g[f"obj{o}_box{b}_rounded"])))
        objects.append((g[f"obj{o}_centers"], g[f"obj{o}_radii"], g[f"obj{o}_pos"], g[f"obj{o}_ori_wxyz"], boxes))
    links = g["self_link_idxs"]
    pairs = [(int(links[a]), int(links[b])) for a, b in g["self_pairs_local"]]
    return build_cost_descriptor(
        pt_link=g["pt_link"], pt_radius=g["pt_radius"], cutoff_obj=float(g["obstacle_cutoff_margin"]),
        objects=objects, ws_min=g["ws_min"], ws_max=g["ws_max"], self_pairs=pairs, self_margin=g["self_margin"],
        ee_link=10 if with_ee else None, ee_target=g["H_target"], ee_w_pos=float(g["ee_w_pos"]),
        ee_w_rot=float(g["ee_w_rot"]), ee_square=bool(g["ee_square"]))


def stretch_cost_from_golden(g):
    from torch_robotics_b200.descriptor import build_cost_descriptor
    return build_cost_descriptor(
        pt_link=g["pt_link"], pt_radius=g["pt_radius"], cutoff_obj=0.01,
        objects=[(g["centers"], g["radii"], None, None)], ws_min=g["ws_min"], ws_max=g["ws_max"],
        ee_link=int(g["ee_link"]), ee_target=g["H_target"])


# ----------------------------------------------------------------------------------------------------------------
# End-to-end decision differences must be PROVEN near-ties (protocol ii).
# The kernel and the oracle evaluate the same decisions (nearest sphere, nearest half-space, margin tests) on link
# positions that differ by a few ulp (different summation order / sin-cos kernels), so a decision may flip only where
# the oracle's own alternatives are within a few ulp of each other.  decision_gaps() recomputes, in float64 on the
# ORACLE's fp32 positions, how far every decision of a configuration is from its boundary.
# ----------------------------------------------------------------------------------------------------------------
TIE_ULPS = 4.0


def decision_gaps(desc, cost, q_rows):
    """Per configuration of q_rows: dict of
       sph_sdf   [n, n_obj, S, Kmax] signed distance of every sphere (inf padding), object pose applied
       ws_d      [n, S, 6]           the six half-space distances in the reference's order
       occ_obj   [n]  min_s |nearest sdf_s - margin_s|     occ_ws [n] min_s |nearest d_s - margin_s|
       occ_self  [n]  min_p |dist_p - margin_p|            scale  [n] magnitude of the operands (positions, centres, >= 1)"""
    from oracle import restatement as orc
    q_rows = np.ascontiguousarray(q_rows, np.float32)
    _, pw = orc.fk_forward(desc, q_rows, np.float32)
    pw = np.asarray(pw, np.float64)
    n = q_rows.shape[0]
    S = len(cost.pt_link)
    x = pw[:, np.asarray(cost.pt_link, np.int64)] if S else np.zeros((n, 0, 3))
    n_obj = len(cost.obj_start) - 1
    kmax = max([int(cost.obj_start[o + 1] - cost.obj_start[o]) for o in range(n_obj)] + [1])
    sph = np.full((n, max(n_obj, 1), S, kmax), np.inf)
    for o in range(n_obj):
        k0, k1 = int(cost.obj_start[o]), int(cost.obj_start[o + 1])
        xo = x
        if not int(cost.obj_identity[o]):
            xo = (x - np.asarray(cost.obj_pos[o], np.float64)) @ np.asarray(cost.obj_R[o], np.float64)
        c = np.asarray(cost.sph_center[k0:k1], np.float64)
        r = np.asarray(cost.sph_radius[k0:k1], np.float64)
        sph[:, o, :, :k1 - k0] = np.linalg.norm(xo[:, :, None, :] - c[None, None], axis=-1) - r
    out = dict(sph_sdf=sph, scale=np.maximum(1.0, np.abs(pw).max(axis=(1, 2))))
    if n_obj:
        out["scale"] = np.maximum(out["scale"], float(np.abs(cost.sph_center).max()) if len(cost.sph_center) else 1.0)
        best = sph.min(axis=-1)                                            # [n, n_obj, S]
        out["occ_obj"] = np.abs(best - np.asarray(cost.pt_margin_obj, np.float64)[None, None]).min(axis=(1, 2)) if S else np.full(n, np.inf)
    else:
        out["occ_obj"] = np.full(n, np.inf)
    if int(cost.ws_enabled) and S:
        d = np.concatenate([x - np.asarray(cost.ws_min, np.float64), np.asarray(cost.ws_max, np.float64) - x], axis=-1)   # [n, S, 6]
        out["ws_d"] = d
        out["occ_ws"] = np.abs(d.min(axis=-1) - np.asarray(cost.pt_margin_ws, np.float64)[None]).min(axis=1)
    else:
        out["ws_d"] = np.full((n, S, 6), np.inf)
        out["occ_ws"] = np.full(n, np.inf)
    P = len(cost.pair_i)
    if P:
        dist = np.linalg.norm(pw[:, np.asarray(cost.pair_i, np.int64)] - pw[:, np.asarray(cost.pair_j, np.int64)], axis=-1)
        out["occ_self"] = np.abs(dist - np.asarray(cost.pair_margin, np.float64)[None]).min(axis=1)
    else:
        out["occ_self"] = np.full(n, np.inf)
    return out


def assert_differences_are_near_ties(desc, cost, q, gpu_argmin, orc_argmin, gpu_occ, orc_occ, grad_flipped_rows, tag, occ_margins=None):
    """Every arg-min / occupancy / gradient-routing difference between the kernel and the oracle must sit within
    TIE_ULPS ulp (of the operands' magnitude) of a decision boundary ON THE ORACLE SIDE.  Prints count + largest gap.
    occ_margins: (self, obj, ws) margin arrays to test the occupancy against when a margin override was used."""
    eps = float(np.finfo(np.float32).eps)
    rows_am = np.unique(np.argwhere(gpu_argmin != orc_argmin)[:, 0]) if gpu_argmin is not None else np.zeros(0, np.int64)
    rows_occ = np.flatnonzero(gpu_occ != orc_occ) if gpu_occ is not None else np.zeros(0, np.int64)
    rows = np.unique(np.concatenate([rows_am, rows_occ, np.asarray(grad_flipped_rows, np.int64)]))
    if rows.size == 0:
        print(f"{tag}: no decision differences")
        return 0
    if occ_margins is not None:
        import copy
        cost = copy.copy(cost)
        cost.pair_margin, cost.pt_margin_obj, cost.pt_margin_ws = occ_margins
    g = decision_gaps(desc, cost, q[rows])
    tol = TIE_ULPS * eps * g["scale"]                       # per row
    pos = {int(r): i for i, r in enumerate(rows)}
    worst = 0.0
    for r, o, s in np.argwhere(gpu_argmin != orc_argmin) if gpu_argmin is not None else []:
        i = pos[int(r)]
        sd = g["sph_sdf"][i, o, s]
        gap = sd[gpu_argmin[r, o, s]] - sd[orc_argmin[r, o, s]]     # how much worse the kernel's choice is on the oracle's positions
        worst = max(worst, abs(gap) / (eps * g["scale"][i]))
        assert abs(gap) <= tol[i], f"{tag}: arg-min of row {r} point {s} differs with a gap of {gap:.3e} (> {TIE_ULPS} ulp)"
    for r in rows_occ:
        i = pos[int(r)]
        diff = int(gpu_occ[r]) ^ int(orc_occ[r])
        for bit, key in ((1, "occ_self"), (2, "occ_obj"), (4, "occ_ws")):
            if diff & bit:
                worst = max(worst, g[key][i] / (eps * g["scale"][i]))
                assert g[key][i] <= tol[i], f"{tag}: occupancy bit {bit} of row {r} differs, distance to the margin {g[key][i]:.3e}"
    for r in np.asarray(grad_flipped_rows, np.int64):
        i = pos[int(r)]
        sd = np.sort(g["sph_sdf"][i], axis=-1)
        gap_sph = (sd[..., 1] - sd[..., 0]).min() if sd.shape[-1] > 1 else np.inf
        best = g["sph_sdf"][i].min(axis=-1)                           # [n_obj, S]: the max over objects routes the gradient too
        so = np.sort(best, axis=0)
        gap_obj = (so[1] - so[0]).min() if so.shape[0] > 1 else np.inf
        wd = np.sort(g["ws_d"][i], axis=-1)
        gap_ws = (wd[:, 1] - wd[:, 0]).min() if wd.size else np.inf
        gap = min(gap_sph, gap_obj, gap_ws)
        worst = max(worst, gap / (eps * g["scale"][i]))
        assert gap <= tol[i], f"{tag}: gradient of row {r} is routed differently but no decision is within {TIE_ULPS} ulp (smallest gap {gap:.3e})"
    print(f"{tag}: {len(rows)} configuration(s) with a decision difference (arg-min rows {len(rows_am)}, occupancy rows {len(rows_occ)}, "
          f"gradient rows {len(grad_flipped_rows)}), all within {worst:.2f} ulp of a decision boundary on the oracle side")
    return len(rows)
