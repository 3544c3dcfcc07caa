#!/bin/bash
# usage: sass_hist.sh <lib.so> <mangled-function-substring>  -- opcode histogram of one kernel
cuobjdump -sass "$1" 2>/dev/null | awk -v f="$2" '
/Function :/ {on = index($0, f) > 0}
on && /^ +\/\*[0-9a-f]+\*\// { op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); n[op]++; tot++ }
END { for (k in n) printf "%6d %s\n", n[k], k; printf "%6d TOTAL\n", tot }' | sort -rn
