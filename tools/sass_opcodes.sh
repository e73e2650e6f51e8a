#!/bin/bash
# SASS evidence that the hot path is tcgen05 / TMEM / TMA (B200_PROFILING.md "What proves a Blackwell-native kernel"):
#   bash tools/sass_opcodes.sh > profiles/rNN_sass_opcodes.txt
SO=vub_image_denoising_b200/libb200dn.so
SASS=$(mktemp)
cuobjdump -sass $SO > $SASS
echo "# cuobjdump -sass $SO : occurrences of each mnemonic (whole-word match; built from HEAD for sm_100a)"
for m in UTCHMMA 'UTCHMMA\.2CTA' LDTM STTM UTMALDG UTMASTG UTCBAR 'UTCBAR\.2CTA\.MULTICAST' UBLKCP HMMA HGMMA QGMMA IGMMA 'SYNCS\.PHASECHK' FADD2 FFMA2 'STG\.E\.128' 'LDG\.E\.128'; do
  printf "%-28s %s\n" "$(echo $m | sed 's/\\//g')" "$(grep -cE "(^|[^A-Z0-9_.])$m([^A-Z0-9_]|$)" $SASS)"
done
echo
echo "# per kernel: UTCHMMA / LDTM / UTMALDG instructions"
awk '/Function :/{name=$3} /UTCHMMA/{u[name]++} /LDTM/{l[name]++} /UTMALDG/{t[name]++} END{for(n in u) printf "%s : %d / %d / %d\n", n, u[n], l[n], t[n]}' $SASS | c++filt \
  | sed -E 's/b200dn::igemm::\(anonymous namespace\):://; s/b200dn::\(anonymous namespace\):://; s/b200dn::igemm:://g' | sort
echo
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor loads, UTCBAR = tcgen05.commit."
echo "# HMMA / HGMMA = 0: no legacy mma.sync / wgmma path.  UTMASTG = 0: the epilogues store with STG (16-byte stores, or 128-byte"
echo "# rows from the shared-memory staged epilogue).  cuBLAS / cuDNN symbols linked into the library:"
nm -D $SO | grep -ci "cublas\|cudnn"
rm -f $SASS
