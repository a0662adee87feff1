#!/usr/bin/env python
"""Opcode mix of a kernel from an ncu report captured with --import-source on:
    python tools/ncu_opcode_mix.py report.ncu-rep [configs_in_the_launch]
Sums 'Instructions Executed' (warp level) per SASS opcode and per class; with the number of configurations of the
launch it also prints warp instructions per configuration (one thread = one configuration)."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    n_cfg = int(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
    per_op = collections.Counter()
    for r in rows[hdr_i + 1:]:
        if len(r) <= ei or not r[ei].strip().isdigit():
            continue
        toks = r[si].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        per_op[op.split(".")[0]] += int(r[ei])
    total = sum(per_op.values())
    classes = collections.OrderedDict([
        ("FP32 FMA pipe (FFMA FMUL FADD)", ("FFMA", "FMUL", "FADD", "FFMA2", "FADD2", "FMUL2")),
        ("FP32 compare / select / min-max (FSETP FSEL FMNMX FMNMX3 FSET)", ("FSETP", "FSEL", "FMNMX", "FMNMX3", "FSET", "FCHK")),
        ("integer / logic / move (IADD3 IMAD LOP3 SHF ISETP SEL MOV LEA ...)", ("IADD3", "IADD", "IMAD", "LOP3", "SHF", "ISETP", "SEL", "MOV", "LEA", "PRMT", "IABS", "PLOP3", "P2R", "R2P", "I2FP", "F2I", "I2F", "F2F", "FRND", "VIADD", "UMOV", "UIADD3", "ULOP3", "USHF", "UISETP", "R2UR", "S2R", "S2UR", "CS2R", "ULDC", "LDC", "UIMAD", "ULEA", "USEL", "UPLOP3", "UFLO", "VOTEU", "VOTE")),
        ("special function (MUFU)", ("MUFU",)),
        ("shared memory (LDS STS)", ("LDS", "STS")),
        ("global / local memory (LDG STG LDL STL)", ("LDG", "STG", "LDL", "STL")),
        ("control (BRA BSSY BSYNC EXIT BAR ...)", ("BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "CALL", "RET", "NOP", "DEPBAR", "BREAK", "BMOV", "YIELD", "NANOSLEEP")),
        ("double precision (sincos range reduction)", ("DFMA", "DMUL", "DADD", "F2F", "DSETP", "FRND")),
    ])
    scale = 1.0 / (n_cfg / 32.0) if n_cfg else None
    print(f"total warp instructions {total}" + (f"  = {total * scale:.0f} per configuration" if scale else ""))
    seen = set()
    for name, ops in classes.items():
        s = sum(per_op[o] for o in ops if o not in seen)
        seen.update(ops)
        print(f"  {100.0 * s / total:5.1f} %  " + (f"{s * scale:7.0f}  " if scale else "") + name)
    rest = {o: c for o, c in per_op.items() if o not in seen}
    s = sum(rest.values())
    print(f"  {100.0 * s / total:5.1f} %  " + (f"{s * scale:7.0f}  " if scale else "") + "other: " + ", ".join(f"{o} {c}" for o, c in sorted(rest.items(), key=lambda kv: -kv[1])[:8]))
    print("top opcodes: " + ", ".join(f"{o} {100.0 * c / total:.1f}%" for o, c in per_op.most_common(14)))


if __name__ == "__main__":
    main()
