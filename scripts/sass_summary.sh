#!/bin/bash
# SASS evidence for the tensor pass: opcode histogram and the tcgen05 / TMA / TMEM / mbarrier instructions of
# tensor_scan_kernel<PAIR=false, EW=8, QRES=true> (the instantiation the 384-d benchmark runs).  CPU only.
# usage: scripts/sass_summary.sh > profiles/r2_k2_tensor_scan_sass.txt
set -eu
LIB=${1:-cortex_b200/libcortex_gpu.so}
TMP=$(mktemp)
cuobjdump -sass "$LIB" | awk '/Function : /{p=($0 ~ /tensor_scan_kernelILb0ELi8ELb1/)} p' > "$TMP"
echo "# $(grep -m1 'Function :' "$TMP")"
echo "# $(grep -c '/\*[0-9a-f]\{4,6\}\*/ ' "$TMP") instructions"
echo
echo "## opcode histogram (top 40)"
grep -o '^\s*/\*[0-9a-f]*\*/\s*\(@!\?U\?P[0-9T]\s\+\)\?[A-Z0-9_.]*' "$TMP" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40
echo
echo "## tensor core (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit), TMA (UTMALDG), TMEM (LDTM = tcgen05.ld, UTCALLOC),"
echo "## mbarrier (SYNCS) and the epilogue's three-input max (FMNMX3)"
grep -E 'UTCHMMA|UTCBAR|UTMALDG|LDTM|UTCALLOC|UTCDEALLOC|FMNMX3|SYNCS' "$TMP" | sed 's/^\s*//; s/\s\+\/\*[0-9a-fx]*\*\/\s*$//' | cut -c1-150
rm -f "$TMP"
