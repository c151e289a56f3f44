#!/bin/bash
# Counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md) per kernel family of the built library.
# usage: bash tools/sass_summary.sh > profiles/<tag>_sass.md
SO=tensorf-jax_b200/tensorf_b200/libtensorf_b200.so
cuobjdump -sass $SO > /tmp/tensorf_sass.txt
echo "# SASS mnemonics in libtensorf_b200.so (cuobjdump -sass, sm_100a)"
echo
echo "| mnemonic | meaning | count |"
echo "|---|---|---:|"
for m in "UTCHMMA:tcgen05.mma (5th-gen tensor cores)" "LDTM:tcgen05.ld (TMEM -> registers)" "STTM:tcgen05.st (registers -> TMEM: A operands of the fused MLP)" \
         "UTCBAR:tcgen05.commit -> mbarrier" "UTMALDG:TMA tiled load (cp.async.bulk.tensor)" "UBLKCP:cp.async.bulk (slab pieces, weights)" \
         "SYNCS:mbarrier arrive / try_wait" "LDGMC:multimem.ld_reduce (NVSwitch in-switch reduction)" "REDG:red.global (scatter-add, vector REDs)"; do
  k=${m%%:*}; d=${m#*:}
  echo "| \`$k\` | $d | $(grep -c "$k" /tmp/tensorf_sass.txt) |"
done
echo
echo "Per kernel (tcgen05.mma / tcgen05.st / cp.async.bulk):"
echo
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /STTM/ {b[name]++} /UBLKCP/ {c[name]++} END {for (n in a) printf "- `%s`: %d UTCHMMA, %d STTM, %d UBLKCP\n", n, a[n], b[n]+0, c[n]+0}' /tmp/tensorf_sass.txt | sed 's/_ZN2tf[0-9]*//' | sort
