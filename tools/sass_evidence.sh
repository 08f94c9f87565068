#!/bin/bash
# Blackwell-native instruction evidence: counts of TMA / mbarrier / named-barrier SASS mnemonics per
# kernel of libswiftly_b200.so (run anywhere nvcc's cuobjdump exists; no GPU needed).
#   tools/sass_evidence.sh > profiles/r02_sass_tma.txt
SO=${1:-ska_sdp_distributed_fourier_transform_b200/libswiftly_b200.so}
echo "# cuobjdump -sass $SO | mnemonic counts per kernel"
echo "# UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, UBLKCP = cp.async.bulk (1-D), UBLKPF = cp.async.bulk.prefetch.L2,"
echo "# SYNCS.* = mbarrier (expect_tx / try_wait), BAR.ARV = bar.arrive (token hand-over of the LSU-token variant, reader hand-over of the TMEM K2),"
echo "# STTM / LDTM = tcgen05.st / tcgen05.ld (tensor memory as parking space), UTCATOMSWS = tcgen05.alloc / dealloc, STG.E...256 = 32-byte stores"
cuobjdump -sass "$SO" 2>/dev/null | awk '
/Function :/ {fn=$3}
{ for (i=1;i<=NF;i++) if ($i ~ /^(UTMALDG|UTMASTG|UBLKCP|UBLKPF|SYNCS|BAR\.ARV|STTM|LDTM|UTCATOMSWS|STG\.E\.[A-Z.0-9]*256)/) { sub(/;$/,"",$i); c[fn"  "$i]++ } }
END { for (k in c) print c[k], k }' | c++filt | sed -E 's/\(swiftly::[^)]*\)//; s/void swiftly::kernel_entry_maps<swiftly:://; s/ >\s/> /' | sort -k2 | awk '{n=$1; $1=""; printf "%5d %s\n", n, $0}'
