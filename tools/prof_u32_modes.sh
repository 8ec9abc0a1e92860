#!/bin/bash
# 32-bit transform kernels (q = 132120577): one polynomial per 4-byte slot (MODE_U32) against two per 8-byte slot (MODE_U32P)
# against the FP64-pipe kernels, forward / inverse / fused product, batch = 2^24 / N polynomials.
for L in 14 13 12 11 10 8; do
  B=$((16777216 >> L))
  echo "--- N=2^$L narrow (one polynomial per 4-byte slot)"; FHEB_U32_PAIR=0 python tools/prof_ntt.py $L 132120577 $B
  echo "--- N=2^$L paired (two polynomials per 8-byte slot)"; FHEB_U32_PAIR=1 python tools/prof_ntt.py $L 132120577 $B
  echo "--- N=2^$L FP64-pipe kernels"; FHEB_NO_U32=1 python tools/prof_ntt.py $L 132120577 $B
done
