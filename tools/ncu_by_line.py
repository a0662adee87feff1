#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counters with nvdisasm line info -> per source line.

usage: python tools/ncu_by_line.py <report.ncu-rep> <library.so | object.o> <kernel-mangled-substring> [n_warps]
(for kernels of the compiled-in family pass the robot's object file, e.g. csrc/build/trk_static_Panda.o: the
per-robot cubins inside libtrk_b200.so share one file name and overwrite each other when extracted)
Needs ncu, cuobjdump and nvdisasm on PATH (no GPU)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, kern = sys.argv[1:4]
    n_warps = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ia, iw = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    counts = [(int(r[ia]), int(r[iw]), r[1].strip()) for r in rows[2:] if len(r) > ia]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    dis = []
    for f in sorted(os.listdir(tmp)):      # the library holds one cubin per translation unit
        if f.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
    lines, on, cur = [], False, ("?", 0)
    for ln in dis:
        if ln.startswith(".text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if len(lines) != len(counts):
        print(f"warning: {len(lines)} disassembled instructions vs {len(counts)} profiled", file=sys.stderr)
    agg, stall = collections.Counter(), collections.Counter()
    for (f, l), (c, w, _) in zip(lines, counts):
        agg[(f, l)] += c
        stall[(f, l)] += w
    tot = sum(agg.values())
    print(f"total warp-instructions {tot}  per warp {tot / n_warps:.1f}")
    src = {}
    for (f, l), c in agg.most_common(70):
        if f not in src:
            libdir = os.path.dirname(os.path.abspath(lib))
            for root in (libdir, os.path.dirname(libdir), "."):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    src[f] = open(p).read().splitlines()
                    break
            else:
                src[f] = []
        text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
        print(f"{c / n_warps:8.1f} {100 * c / tot:5.1f}% stall {stall[(f, l)]:6d}  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
