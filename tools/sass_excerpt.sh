#!/bin/bash
# Regenerates profiles/r02_sass_excerpt.md: the tensor-core / TMEM / packed-fp32 instructions of the shipped kernels, straight
# from libgode.so (cuobjdump -sass), so that "runs on tcgen05" can be checked without a GPU.
cd "$(dirname "$0")/.." || exit 1
SO=graph-odenet_b200/csrc/libgode.so
OUT=profiles/r02_sass_excerpt.md
{
echo "# r02 — SASS evidence (cuobjdump -sass $SO, sm_100a)"
echo
echo "Mnemonics per kernel (counts of static instructions): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers),"
echo "UTCBAR = tcgen05.commit, FADD2 / FFMA2 = packed fp32 (add.f32x2 / fma.f32x2)."
echo
echo '| kernel | UTCHMMA | LDTM | UTCBAR | FADD2 | FFMA2 |'
echo '|---|---:|---:|---:|---:|---:|'
cuobjdump -sass $SO 2>/dev/null | awk '
/Function :/ {f=$3}
/UTCHMMA/ {a[f]++} /LDTM/ {b[f]++} /UTCBAR/ {c[f]++} /FADD2/ {d[f]++} /FFMA2/ {e[f]++}
END {for (k in a) s[k]=1; for (k in d) s[k]=1; for (k in e) s[k]=1;
     for (k in s) printf "| `%s` | %d | %d | %d | %d | %d |\n", k, a[k], b[k], c[k], d[k], e[k]}' | sort | while read -r line; do
  sym=$(echo "$line" | sed -E 's/^\| `([^`]*)`.*/\1/'); dem=$(echo "$sym" | c++filt | sed -E 's/\(.*//; s/^void //')
  echo "$line" | sed "s|\`$sym\`|\`$dem\`|"
done
echo
echo "## Excerpt: the MMA issue loop of the transform (\`k_rows_ws<128,4,0,19>\`)"
echo
echo '```'
ROWS_WS=$(cuobjdump -sass $SO 2>/dev/null | grep "Function :" | grep -o "_ZN4gode9k_rows_wsILi128ELi4ELi0ELi19E[A-Za-z0-9_]*" | head -1)   # the mangled name follows the signature
cuobjdump -sass -fun "$ROWS_WS" $SO 2>/dev/null | grep -E "UTCHMMA|UTCBAR|LDTM|SYNCS" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/^\s+//' | head -28
echo '```'
echo
echo "## Excerpt: the inner loop of the gather (\`k_spmm_t2<32,8,true>\`): 4 LDS + 4 IMAD.WIDE.U32 + 4 LDG.E.128 + 8 FADD2 per four neighbour rows"
echo
echo '```'
T2=$(cuobjdump -sass $SO 2>/dev/null | grep "Function :" | grep -o "_ZN4gode9k_spmm_t2ILi32ELi8ELb1E[A-Za-z0-9_]*" | head -1)
cuobjdump -sass -fun "$T2" $SO 2>/dev/null | grep -E "/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/^\s+//' | awk '/LDS R[0-9]+, \[R[0-9]+\] ;/ && !p {p=1} p {print; n++} n>=26 {exit}'
echo '```'
} > $OUT
wc -l $OUT
