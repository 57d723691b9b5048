#!/bin/bash
# Tensor-core / TMA / TMEM mnemonics per kernel of the built library (evidence for profiles/):
#   tools/sass_extract.sh > profiles/r02_sass_tensor_mnemonics.txt
LIB=${1:-fastselect_b200/lib/libfastselect_b200.so}
echo "# cuobjdump -sass $LIB: tensor-core / TMA / TMEM mnemonics per kernel (count, kernel, mnemonic)"
echo "# UTCOMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tile load (.4D / .5D: the permuted operand boxes of the"
echo "# merged accumulation kernel), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (16dp256bit = the 16x256b fragment)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3 }
  /UTCOMMA|UTMALDG|UTCBAR|LDTM|STTM|UTMAPF/ {
    for (i = 1; i <= NF; ++i) if ($i ~ /^(UTCOMMA|UTMALDG|UTCBAR|LDTM|STTM|UTMAPF)/) { m=$i; break }
    c[fn " :: " m]++
  }
  END { for (k in c) print c[k] "\t" k }' | c++filt | sed -E "s/\(CUtensorMap_st.*\) :: / :: /" | sort -k2
