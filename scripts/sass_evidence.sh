#!/bin/bash
# SASS evidence that K1 stages its record tiles with TMA bulk copies on mbarriers (runs without a GPU):
#   UBLKCP.S.G = cp.async.bulk global -> shared, SYNCS.* = mbarrier ops, ELECT = elect.sync.
OUT=${1:-profiles/r02/sass_k1_tma.txt}
SO=besst_b200/libbesst_b200.so
{
  echo "# cuobjdump -sass $SO  (built by besst_b200/build.py: nvcc -gencode arch=compute_100a,code=sm_100a)"
  echo "# per kernel: count of TMA bulk copies (UBLKCP), mbarrier ops (SYNCS.*), elect.sync (ELECT)"
  cuobjdump -sass $SO 2>/dev/null | grep -E "^\s+Function : |UBLKCP|SYNCS\.|ELECT|arch = " \
    | awk '/arch =/{a=$3} /Function/{f=$3; arch[f]=a} /UBLKCP/{u[f]++} /SYNCS/{s[f]++} /ELECT/{e[f]++} END{for(k in u) print arch[k], k, "UBLKCP", u[k], "SYNCS", s[k], "ELECT", e[k]}' | sort
  echo
  echo "# k_extract_links_tma<INT_RL=true, PACKED=true>: the TMA / mbarrier instructions in address order"
  FN=$(cuobjdump -sass $SO 2>/dev/null | grep -oE "_ZN[0-9A-Za-z_]*k_extract_links_tmaILb1ELb1E[0-9A-Za-z_]*" | head -1)
  cuobjdump -sass -fun "$FN" $SO 2>/dev/null \
    | grep -E "UBLKCP|SYNCS|ELECT" | sed 's/ *\/\* 0x[0-9a-f]* \*\///'
  echo
  echo "# all cubins in the library:"
  cuobjdump -lelf $SO 2>/dev/null
} > $OUT
wc -l $OUT
