#!/bin/bash
# Counts the SASS mnemonics that prove which hardware paths the built library uses (tcgen05 MMA / TMEM loads / bulk copies).
# usage: tools/sass_evidence.sh > profiles/r02_sass_mnemonics.txt
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
so=$here/msolve_b200/csrc/libf4la_b200.so
tmp=$(mktemp)
cuobjdump -sass "$so" > "$tmp"
echo "cuobjdump -sass msolve_b200/csrc/libf4la_b200.so (sm_100a), $(grep -c 'Function :' "$tmp") kernels"
for m in UTCIMMA UTCBAR UTCATOMSWS LDTM STTM UBLKCP SYNCS IMAD.WIDE.U32 ATOMG; do
    printf "%-14s %6d\n" "$m" "$(grep -c "$m" "$tmp")"
done
echo "UTCIMMA (tcgen05.mma.kind::i8) per kernel:"
awk '/Function :/{f=$3} /UTCIMMA/{c[f]++} END{for(k in c) printf "  %-50s %d\n", k, c[k]}' "$tmp"
echo "LDTM (tcgen05.ld) per kernel:"
awk '/Function :/{f=$3} /LDTM/{c[f]++} END{for(k in c) printf "  %-50s %d\n", k, c[k]}' "$tmp"
rm -f "$tmp"
