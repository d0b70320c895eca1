#!/bin/bash
# per-role cycle shares of conv_tc.cu (SAUNET_TC_PROF counters) on the step's 1x1 / conv-transpose-phase shapes
for spec in "16 32 32 512 128 1 PRO" "16 16 16 768 128 1 PRO" "16 64 64 128 128 2" "16 16 16 512 512 2" "16 128 128 64 128 1 PRO" "16 64 64 256 128 1 PRO"; do
  set -- $spec
  for prec in ${PRECS:-3xtf32 bf16}; do
    if [ "$7" = PRO ]; then export PRO=1; else unset PRO; fi
    echo "== $prec ${7:-raw} $1 $2 $3 $4 $5 k$6"; PROF=1 SAUNET_CONV1X1_T=0 SAUNET_PRECISION=$prec timeout 60 python tools/bench_conv.py fwd $1 $2 $3 $4 $5 $6 20 2>&1 | tail -2
  done
done
