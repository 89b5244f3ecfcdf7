#!/bin/bash
# usage: tools/sass_census.sh <mangled-kernel-name-substring>   -- opcode census of one kernel
LIB=$(dirname "$0")/../bioshell_b200/libbioshell_align.so
cuobjdump -sass "$LIB" | awk -v pat="$1" '
  /Function :/ { on = index($0, pat) > 0; if (on) print $0 }
  on && /^ +\/\*[0-9a-f]+\*\/ +[@A-Z]/ { op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op); c[op]++; n++ }
  END { for (k in c) printf "%6d %s\n", c[k], k | "sort -rn"; close("sort -rn"); print n " instructions" }'
