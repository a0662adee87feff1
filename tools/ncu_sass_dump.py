#!/usr/bin/env python
"""Instruction-level listing of a kernel from an ncu report captured with --import-source on:
    python tools/ncu_sass_dump.py report.ncu-rep [n_warps] > listing.txt
One line per SASS instruction: executed count per warp, stall samples (total | math dispatch barrier wait short_sb
long_sb no_inst not_selected), the instruction.  Followed by totals of the stall columns."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    n_warps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    col = {n: h.index(n) for n in ("Source", "Instructions Executed", "Warp Stall Sampling (All Samples)", "stall_math", "stall_dispatch",
                                   "stall_barrier", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_no_inst", "stall_not_selected",
                                   "stall_selected", "stall_branch_resolving", "stall_mio", "stall_lg")}
    tot = {k: 0 for k in col if k.startswith("stall_") or k.startswith("Warp")}
    n_inst = 0
    for r in rows[hi + 1:]:
        if len(r) <= col["stall_lg"]:
            continue
        ex = int(r[col["Instructions Executed"]])
        n_inst += ex
        vals = {k: int(r[c]) for k, c in col.items() if k != "Source" and k != "Instructions Executed"}
        for k in tot:
            tot[k] += vals[k]
        print(f"{ex / n_warps:7.2f} {vals['Warp Stall Sampling (All Samples)']:6d} | m{vals['stall_math']:5d} d{vals['stall_dispatch']:5d} b{vals['stall_barrier']:5d} "
              f"w{vals['stall_wait']:5d} s{vals['stall_short_sb']:5d} l{vals['stall_long_sb']:5d} n{vals['stall_no_inst']:5d} ns{vals['stall_not_selected']:5d} | {r[col['Source']].strip()}")
    print("executed per warp", n_inst / n_warps)
    print("stall sample totals", tot)


if __name__ == "__main__":
    main()
