#!/usr/bin/env python
"""Emit torch_robotics_b200/csrc/robots_generated.cuh: constexpr kinematic tables of the robots the
package ships, consumed by the register-resident kernel family (trk_static.cuh).

Each robot becomes a struct of `static constexpr` arrays in parents-first order.  Floats are written as
hexadecimal literals so the compiled-in constants are bit-identical to what compile_urdf() produces
(and therefore to the reference's fp32 constants).  The struct also carries the exact word image of
the runtime tree buffer (trk_device.cuh: TreeHeader + LinkRec[L]); trk_model_create compares a model
against these images to decide whether a compiled-in family can serve it.

usage: python tools/gen_static_robots.py        (rewrites the header in place)
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_robotics_b200.descriptor import compile_urdf  # noqa: E402

DATA = os.path.join(ROOT, "torch_robotics_b200", "data", "urdf", "robots")
ROBOTS = [
    ("panda", "franka_description/robots/panda_arm_no_gripper.urdf"),
    ("panda_gripper", "franka_description/robots/panda_arm_hand.urdf"),
    ("stretch", "habitat_stretch/urdf/hab_stretch.urdf"),
    ("ur10", "ur10/urdf/ur10.urdf"),
    ("kuka_iiwa", "kuka_iiwa/urdf/iiwa7.urdf"),
    ("tiago_dual_holo", "tiago_dual_description/tiago_dual_holobase_minimal.urdf"),
    ("tiago_dual_holo_move", "tiago_dual_description/tiago_dual_holobase_minimal_holonomic.urdf"),
    ("shadow_hand", "shadow_hand/shadow_hand.urdf"),
    ("allegro_hand", "allegro_hand/allegro_hand.urdf"),
    ("planar_2link", "planar_manipulators/urdf/2_link_planar.urdf"),
]


def hexf(x):
    x = float(np.float32(x))
    if x == 0.0:
        return "-0.0f" if np.signbit(np.float32(x)) else "0.0f"
    return float.hex(x) + "f"


def f2u(x):
    return struct.unpack("<I", struct.pack("<f", float(np.float32(x))))[0]


def tree_words(d):
    """Word image of the runtime tree buffer - must mirror trk_model_create() in trk_b200.cu."""
    L = d.n_links
    w = [0] * (16 + 24 * L)
    w[0], w[1] = L, d.n_dofs
    ident = [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]
    for i, v in enumerate(ident):
        w[4 + i] = f2u(v)
    for i in range(L):
        b = int(d.topo[i])
        o = 16 + 24 * i
        for k in range(9):
            w[o + k] = f2u(d.R_fixed[b].reshape(9)[k])
        for k in range(3):
            w[o + 9 + k] = f2u(d.trans[b][k])
            w[o + 12 + k] = f2u(d.axis[b][k])
        w[o + 15] = f2u(d.axis_s[b])
        w[o + 16], w[o + 17] = f2u(d.lo[b]), f2u(d.hi[b])
        w[o + 18] = int(d.kind[b]) | (int(d.axis_k[b]) << 2) | (16 if d.has_lim[b] else 0)
        w[o + 19] = int(d.parent[b]) & 0xFFFFFFFF
        w[o + 20] = b
        w[o + 21] = int(d.q_index[b]) & 0xFFFFFFFF
    return w


def emit_robot(name, d):
    L = d.n_links
    order = [int(b) for b in d.topo]
    cname = "".join(p.capitalize() for p in name.split("_"))
    out = [f"struct {cname} {{",
           f'    static constexpr const char* name = "{name}";',
           f"    static constexpr int L = {L}, D = {d.n_dofs};"]

    def iarr(nm, vals):
        out.append(f"    static constexpr int {nm}[L] = {{{', '.join(str(int(v)) for v in vals)}}};")

    def farr(nm, vals):
        out.append(f"    static constexpr float {nm}[L] = {{{', '.join(hexf(v) for v in vals)}}};")

    iarr("body", order)
    iarr("parent", [d.parent[b] for b in order])
    iarr("kind", [d.kind[b] for b in order])
    iarr("axis_k", [d.axis_k[b] for b in order])
    iarr("qidx", [d.q_index[b] for b in order])
    iarr("has_lim", [d.has_lim[b] for b in order])
    farr("axis_s", [d.axis_s[b] for b in order])
    farr("lo", [d.lo[b] for b in order])
    farr("hi", [d.hi[b] for b in order])
    for nm, arr, n in (("Rf", d.R_fixed.reshape(L, 9), 9), ("t", d.trans, 3), ("ax", d.axis, 3)):
        rows = ["{" + ", ".join(hexf(v) for v in arr[b]) + "}" for b in order]
        out.append(f"    static constexpr float {nm}[L][{n}] = {{\n        " + ",\n        ".join(rows) + "};")
    words = tree_words(d)
    out.append(f"    static constexpr int n_words = {len(words)};")
    lines = [", ".join(f"0x{v:08x}u" for v in words[i:i + 8]) for i in range(0, len(words), 8)]
    out.append("    static constexpr unsigned words[n_words] = {\n        " + ",\n        ".join(lines) + "};")
    out.append("};")
    return cname, "\n".join(out)


def emit_panda_collision():
    """Compile-time collision model of RobotPanda (robots/robot_panda.py:47-127 in the reference; constants in
    torch_robotics_b200/robots/robot_panda.py): collision-point bodies and self-collision body pairs."""
    from torch_robotics_b200.robots.robot_panda import OBJECT_COLLISION_SPHERES, SELF_COLLISION_PAIRS, GRASPED_OBJECT_SELF_LINKS
    d = compile_urdf(os.path.join(DATA, dict(ROBOTS)["panda"]), "panda")
    idx = d.name_to_idx()
    pts = [idx[n] for n in OBJECT_COLLISION_SPHERES]
    links = set(GRASPED_OBJECT_SELF_LINKS)
    for k, v in SELF_COLLISION_PAIRS.items():
        links.add(k)
        links.update(v)
    links = sorted(links)
    pairs = [(idx[a], idx[b]) for a in links for b in SELF_COLLISION_PAIRS.get(a, [])]
    out = ["struct PandaCollision {   // RobotPanda: 5 object-collision points, 10 self-collision pairs",
           f"    static constexpr int S = {len(pts)}, P = {len(pairs)};",
           f"    static constexpr int pt_link[S] = {{{', '.join(map(str, pts))}}};",
           f"    static constexpr int pair_i[P] = {{{', '.join(str(a) for a, _ in pairs)}}};",
           f"    static constexpr int pair_j[P] = {{{', '.join(str(b) for _, b in pairs)}}};",
           f"    static constexpr int ee_body = {idx['ee_link']};   // end-effector body of the planning objective (RobotPanda.link_name_ee)",
           "};"]
    return "\n".join(out)


def emit_stretch_collision():
    """Compile-time collision model of BASELINE config 5 (the reference has no Stretch
    collision model): collision points at the origins of the 14 movable links + link_straight_gripper, no self pairs."""
    d = compile_urdf(os.path.join(DATA, dict(ROBOTS)["stretch"]), "stretch")
    pts = list(d.controlled) + [d.link_names.index("link_straight_gripper")]
    return "\n".join(["struct StretchCollision {   // config 5: 15 object / workspace collision points, no self-collision pairs",
                      f"    static constexpr int S = {len(pts)}, P = 0;",
                      f"    static constexpr int pt_link[S] = {{{', '.join(map(str, pts))}}};",
                      "    static constexpr int pair_i[1] = {0};",
                      "    static constexpr int pair_j[1] = {0};",
                      f"    static constexpr int ee_body = {d.link_names.index('link_straight_gripper')};",
                      "};"])


def main():
    parts = ["// GENERATED by tools/gen_static_robots.py - do not edit.",
             "// constexpr kinematic tables of the shipped robots (parents-first order).",
             "#pragma once", "", "namespace trk {", "namespace robots {", ""]
    names = []
    for name, rel in ROBOTS:
        d = compile_urdf(os.path.join(DATA, rel), name)
        cname, text = emit_robot(name, d)
        names.append(cname)
        parts += [text, ""]
    parts += [emit_panda_collision(), "", emit_stretch_collision(), ""]
    parts += ["}  // namespace robots", "",
              "#define TRK_FOR_EACH_STATIC_ROBOT(X) " + " ".join(f"X(robots::{n})" for n in names), "",
              "}  // namespace trk", ""]
    path = os.path.join(ROOT, "torch_robotics_b200", "csrc", "robots_generated.cuh")
    with open(path, "w") as f:
        f.write("\n".join(parts))
    print("wrote", path, "robots:", ", ".join(names))


if __name__ == "__main__":
    main()
