#!/bin/bash
# tools/sass_evidence.sh -- instruction-mix excerpts of the hot kernels from the built library (no GPU needed):
# which memory / synchronisation instructions each kernel was compiled to. Output: profiles/r02_f_sass_evidence.md
cd "$(dirname "$0")/.."
LIB=mygram-db_b200/libmgx.so
OUT=profiles/r02_f_sass_evidence.md
{
echo "# SASS evidence (\`cuobjdump -sass $LIB\`, sm_100a), round 2"
echo
echo "Produced by \`tools/sass_evidence.sh\` from the committed sources (nvcc 12.9, \`-gencode arch=compute_100a,code=sm_100a\`)."
echo "Per kernel: registers / shared memory from \`cuobjdump --dump-resource-usage\`, then the count of every memory,"
echo "synchronisation and warp-collective instruction in its SASS."
for k in radix_onesweep_kernel radix_scatter_kernel tokenize_flat_kernelILi2E df_units_kernel df_tile_kernel and_tiles_kernel and_tile_kernel df_stream_kernel topk_kernel merge_topk_kernel; do
  fn=$(cuobjdump -sass $LIB 2>/dev/null | grep "Function :" | grep "$k" | head -1 | sed 's/.*Function : //')
  [ -z "$fn" ] && continue
  echo
  echo "## $k"
  echo
  cuobjdump --dump-resource-usage $LIB 2>/dev/null | grep -A1 "Function $fn" | grep -o "REG:[0-9]*\|SHARED:[0-9]*\|STACK:[0-9]*" | paste -sd' '
  echo '```'
  cuobjdump -sass -fun "$fn" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*$//; s/^@!?U?P[0-9T] //' | awk '{print $1}' \
    | grep -E "^(LDG|STG|LDS|STS|LDSM|ATOM|ATOMS|ATOMG|RED|MATCH|SHFL|VOTE|BAR|SYNCS|UBLKCP|UTMA|FENCE|MEMBAR|NANOSLEEP|LDC|LDL|STL|DFMA|DMUL|DADD|MUFU|CCTL|ERRBAR|WARPSYNC|REDUX|POPC|DSETP)" \
    | sort | uniq -c | sort -rn | head -28
  echo '```'
done
} > $OUT
wc -l $OUT
