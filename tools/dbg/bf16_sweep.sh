#!/bin/bash
# 3x3 kernel: 3xTF32 vs bf16 operands, with / without BatchNorm statistics in the epilogue (what bounds the TMA kernel?)
for spec in "16 256 256 64 64" "16 256 256 32 32" "16 128 128 128 32 PRO" "16 64 64 128 32 PRO" "16 128 128 32 128" "16 64 64 256 128"; do
  set -- $spec
  for prec in ${PRECS:-3xtf32 bf16}; do
    for ns in ${NOSTATS:-"" 1}; do
      if [ "$6" = PRO ]; then export PRO=1; else unset PRO; fi
      if [ -n "$ns" ]; then export NOSTAT=1; else unset NOSTAT; fi
      echo -n "$prec ${6:-raw} nostat=${ns:-0} "; SAUNET_PRECISION=$prec python tools/bench_conv.py fwd $1 $2 $3 $4 $5 3 20 2>&1 | tail -1
    done
  done
done
