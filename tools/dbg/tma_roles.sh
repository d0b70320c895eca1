#!/bin/bash
# per-role cycle shares of the TMA-fed 3x3 kernel (SAUNET_TC_PROF counters) on the step's shapes, both operand classes
for spec in "16 256 256 64 64" "16 128 128 128 32 PRO" "16 64 64 128 32 PRO" "16 128 128 32 128" "16 64 64 256 128"; do
  set -- $spec
  for prec in ${PRECS:-3xtf32 bf16}; do
    if [ "$6" = PRO ]; then export PRO=1; else unset PRO; fi
    echo "== $prec ${6:-raw} $1 $2 $3 $4 $5"; PROF=tma SAUNET_PRECISION=$prec timeout 60 python tools/bench_conv.py fwd $1 $2 $3 $4 $5 3 20 2>&1 | tail -2
  done
done
