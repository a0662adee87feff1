#!/usr/bin/env python
"""Static instruction count per source line of one kernel, from `nvdisasm -g -c <cubin>` output:
    cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > dis.txt; python tools/sass_by_line.py dis.txt <kernel-name-substring> [top]
Counts SASS instructions under each `//## File "...", line N` marker (innermost inlined location)."""
import collections
import re
import sys


def main():
    txt = open(sys.argv[1]).read().split("\n")
    pat = sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    on = False
    cur = ("?", 0)
    per = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    total = 0
    for ln in txt:
        if ln.startswith("//---") and ".text." in ln:
            on = pat in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            per[cur] += 1
            ops[cur][m.group(2).split(".")[0]] += 1
            total += 1
    print("static instructions", total)
    src_cache = {}
    for (f, l), c in per.most_common(top):
        print(f"{c:5d}  {f}:{l}  " + " ".join(f"{o}{n}" for o, n in ops[(f, l)].most_common(6)))


if __name__ == "__main__":
    main()
