"""Compile-time specialisation of the CUDA library for a robot.

The fast kernels of csrc/ (trk_static.cuh, trk_hot.cuh, trk_ext.cuh) are templates over a struct of `static constexpr`
kinematic tables: the tree is unrolled, fixed chains are folded, structural zeros vanish.  The ten shipped robots are emitted
into csrc/robots_generated.cuh at build time (tools/gen_static_robots.py uses the functions below); any OTHER descriptor - a
user's URDF - is served by the table-driven runtime-tree kernels, which are 3 - 20x slower.

`build_specialized_library(desc, collision)` closes that gap at run time: it writes the same kind of header for the one robot
(`robots::Jit`, optionally its collision model `robots::JitCollision`), compiles a private copy of the library against it
with nvcc (`-DTRK_ROBOTS_HEADER=...`, two translation units in parallel, ~1 minute) and caches the result by content hash.
The private library exports the same C ABI (include/trk_b200.h); `DeviceModel(desc, device, library=...)` binds a model to it.
`DifferentiableTree.specialize()` is the user-facing call.

The reference has no counterpart (it interprets the tree in PyTorch, robot_tree.py:135-216); the specialised kernels are the
compiled-in family, so their parity with the reference is that family's (tests/test_specialize.py checks a non-shipped
robot against the oracle and against the runtime-tree kernels).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import struct
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
MAX_HOT_POINTS = 16      # kHotMaxPoints / kHotMaxPairs of trk_hot.cuh
MAX_HOT_PAIRS = 16


def hexf(x) -> str:
    """fp32 value as a hexadecimal literal: the compiled-in constant is bit-identical to the descriptor's."""
    x = float(np.float32(x))
    if x == 0.0:
        return "-0.0f" if np.signbit(np.float32(x)) else "0.0f"
    return float.hex(x) + "f"


def f2u(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(np.float32(x))))[0]


def tree_words(d):
    """Word image of the runtime tree buffer - must mirror trk_model_create() in trk_b200.cu (TreeHeader + LinkRec[L]):
    trk_model_create compares a model against these images to decide whether a compiled-in family can serve it."""
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


def emit_robot(name: str, d, cname: Optional[str] = None) -> Tuple[str, str]:
    """One robot as a struct of constexpr tables in parents-first order -> (struct name, source text)."""
    L = d.n_links
    order = [int(b) for b in d.topo]
    cname = cname or "".join(p.capitalize() for p in name.split("_"))
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


def emit_collision(cname: str, pt_link: Sequence[int], pairs: Sequence[Tuple[int, int]], ee_body: int, comment: str = "") -> str:
    """A compile-time collision model: bodies of the collision points, self-collision body pairs, end-effector body."""
    pi = [int(a) for a, _ in pairs] or [0]
    pj = [int(b) for _, b in pairs] or [0]
    return "\n".join([f"struct {cname} {{" + (f"   // {comment}" if comment else ""),
                      f"    static constexpr int S = {len(pt_link)}, P = {len(pairs)};",
                      f"    static constexpr int pt_link[{max(len(pt_link), 1)}] = {{{', '.join(str(int(v)) for v in pt_link) or '0'}}};",
                      f"    static constexpr int pair_i[{len(pi)}] = {{{', '.join(map(str, pi))}}};",
                      f"    static constexpr int pair_j[{len(pj)}] = {{{', '.join(map(str, pj))}}};",
                      f"    static constexpr int ee_body = {int(ee_body)};",
                      "};"])


def specialized_header(desc, collision: Optional[dict] = None) -> str:
    """Header of a one-robot library.  collision: {"pt_link": [...], "pairs": [(i, j), ...], "ee_body": b} (body indices) or None."""
    if not np.allclose(np.asarray(desc.base_R, np.float32), np.eye(3)) or np.any(np.asarray(desc.base_t, np.float32) != 0):
        raise ValueError("a specialised library is compiled for the robot in its own base frame: reset the base pose first")
    if desc.n_links < 2:
        raise ValueError("nothing to specialise: the tree has a single body")
    # The struct names carry a content hash: two specialised libraries in one process must not share symbols.  (Inline
    # variables such as Jit::words are STB_GNU_UNIQUE: the dynamic loader would resolve the second library's table to the
    # first one's, whatever RTLD_LOCAL says.  The build also passes -fno-gnu-unique and links with -Bsymbolic.)
    _, probe = emit_robot(getattr(desc, "name", None) or "jit", desc, cname="Jit")
    tag = hashlib.sha256((probe + repr(collision)).encode()).hexdigest()[:12]
    _, text = emit_robot(getattr(desc, "name", None) or "jit", desc, cname=f"Jit_{tag}")
    parts = ["// GENERATED by torch_robotics_b200/specialize.py - one robot, compiled at run time.", "#pragma once", "",
             "namespace trk {", "namespace robots {", "", text, f"using Jit = Jit_{tag};", ""]
    if collision is not None:
        S, P = len(collision["pt_link"]), len(collision["pairs"])
        if S > MAX_HOT_POINTS or P > MAX_HOT_PAIRS:
            raise ValueError(f"compile-time collision models hold <= {MAX_HOT_POINTS} points and <= {MAX_HOT_PAIRS} pairs "
                             f"(got {S}, {P}); specialise the robot alone (collision=None) and keep the table-driven cost kernels")
        parts += [emit_collision(f"JitCollision_{tag}", collision["pt_link"], collision["pairs"], collision["ee_body"]),
                  f"using JitCollision = JitCollision_{tag};", ""]
    parts += ["}  // namespace robots", ""]
    if collision is not None:
        parts += ["#define TRK_JIT_HAS_COLLISION 1"]
    parts += ["#define TRK_FOR_EACH_STATIC_ROBOT(X) X(robots::Jit)", "", "}  // namespace trk", ""]
    return "\n".join(parts)


def _sources_digest() -> str:
    h = hashlib.sha256()
    for root, names in ((CSRC, sorted(os.listdir(CSRC))), (INCLUDE, sorted(os.listdir(INCLUDE)) if os.path.isdir(INCLUDE) else [])):
        for f in names:
            if f.endswith((".cu", ".cuh", ".h")) and f != "robots_generated.cuh":
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def default_cache_dir() -> str:
    return os.environ.get("TRK_B200_JIT_DIR") or os.path.join(CSRC, "jit")


def library_path(desc, collision: Optional[dict] = None, cache_dir: Optional[str] = None) -> str:
    """Where the specialised library of (descriptor, collision model, current sources) lives, built or not."""
    key = hashlib.sha256((specialized_header(desc, collision) + _sources_digest()).encode()).hexdigest()[:20]
    return os.path.join(cache_dir or default_cache_dir(), f"libtrk_jit_{key}.so")


def build_specialized_library(desc, collision: Optional[dict] = None, cache_dir: Optional[str] = None, verbose: bool = False) -> str:
    """Compile (or find in the cache) the private library of one robot -> path of the .so.

    Needs nvcc (env NVCC, else /usr/local/cuda/bin/nvcc, else PATH); cross-compiles for sm_100a, so it also runs on a
    machine without a GPU.  Raises RuntimeError with the compiler's output when the build fails."""
    out = library_path(desc, collision, cache_dir)
    if os.path.exists(out):
        return out
    nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else shutil.which("nvcc"))
    if not nvcc:
        raise RuntimeError("specialising a robot needs nvcc (set NVCC or put it on PATH); without it the runtime-tree kernels serve the model")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    work = tempfile.mkdtemp(prefix="trk_jit_")
    try:
        hdr = os.path.join(work, "robots_jit.cuh")
        with open(hdr, "w") as f:
            f.write(specialized_header(desc, collision))
        common = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                  "-Xcompiler", "-fno-gnu-unique", f'-DTRK_ROBOTS_HEADER="{hdr}"', "-I", CSRC, "-c"]
        jobs = [common + ["-o", os.path.join(work, "trk_b200.o"), os.path.join(CSRC, "trk_b200.cu")],
                common + ["-DTRK_INST_ROBOT=robots::Jit", "-o", os.path.join(work, "trk_static_jit.o"), os.path.join(CSRC, "trk_static_inst.cu")]]
        with ThreadPoolExecutor(max_workers=2) as ex:
            for r in ex.map(lambda c: subprocess.run(c, cwd=CSRC, capture_output=True, text=True), jobs):
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed while specialising the robot:\n" + r.stderr[-4000:])
                if verbose and r.stderr:
                    print(r.stderr[-2000:])
        tmp_out = os.path.join(work, "libtrk_jit.so")
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xlinker", "-Bsymbolic", "-o", tmp_out, jobs[0][jobs[0].index("-o") + 1],
                            jobs[1][jobs[1].index("-o") + 1]], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed while linking the specialised library:\n" + r.stderr[-4000:])
        staged = out + f".tmp{os.getpid()}"        # same directory as the target: the final rename is atomic, a concurrent builder
        shutil.move(tmp_out, staged)               # of the same robot (another rank) never sees a half-written library
        os.replace(staged, out)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return out
