#!/bin/bash
# static opcode histogram of one kernel in an object / shared library:  tools/sass_mix.sh <file> <mangled-name-substring>
f=$1; pat=$2
cuobjdump -sass "$f" | awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0} on {print}' > /tmp/_k.sass
n=$(grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/_k.sass)
echo "static instructions: $n  ($((n*16)) bytes)"
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/_k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+).*/\2/' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-40}
