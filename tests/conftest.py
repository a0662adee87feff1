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
