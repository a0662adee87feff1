"""Run-time specialised libraries the tests (and tools/bench_generic.py) use: built by __graft_entry__.build() in the build
container so that they travel to the GPU box with the snapshot; the tests build them on demand otherwise (nvcc, ~1 minute)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "torch_robotics_b200", "data", "urdf", "robots")


def iiwa7_allegro_desc():
    """A robot outside the shipped set: 29 links, 23 DoF (the iiwa7 arm + the Allegro hand's four 4-joint fingers)."""
    from torch_robotics_b200.descriptor import compile_urdf
    return compile_urdf(os.path.join(DATA, "kuka_iiwa", "urdf", "iiwa7_allegro.urdf"), "iiwa7_allegro")


def ur10_collision():
    """An ad-hoc collision model for the UR10 (compiled in as a robot, but without a compile-time collision model):
    five collision points, four self-collision pairs, the tool link as the end-effector."""
    from conftest import get_desc
    d = get_desc("ur10")
    idx = d.name_to_idx()
    pts = [idx[n] for n in ("shoulder_link", "upper_arm_link", "forearm_link", "wrist_1_link", "wrist_3_link")]
    pairs = [(idx["wrist_3_link"], idx["base_link"]), (idx["wrist_3_link"], idx["shoulder_link"]),
             (idx["wrist_1_link"], idx["base_link"]), (idx["forearm_link"], idx["base_link"])]
    return d, dict(pt_link=pts, pairs=pairs, ee_body=d.n_links - 1)


def build_all(verbose=False):
    from concurrent.futures import ThreadPoolExecutor
    from torch_robotics_b200 import specialize as sp
    d, cm = ur10_collision()
    with ThreadPoolExecutor(max_workers=2) as ex:
        return list(ex.map(lambda a: sp.build_specialized_library(a[0], a[1], verbose=verbose), [(iiwa7_allegro_desc(), None), (d, cm)]))
