#!/bin/bash
# compare the TMA-fed 3x3 kernel with the LDG-gather kernels on the shapes of the step
for spec in "16 128 128 128 32 PRO" "16 64 64 128 32 PRO" "16 32 32 128 32 PRO" "16 16 16 128 32 PRO" "16 256 256 64 64" "16 256 256 32 32" "16 256 256 16 16" "16 256 256 64 32" "16 128 128 256 64" "16 128 128 32 128" "16 32 32 32 128"; do
  set -- $spec
  for mode in tma old; do
    if [ "$mode" = old ]; then export SAUNET_NO_HALO_TMA=1; else unset SAUNET_NO_HALO_TMA; fi
    if [ "$6" = PRO ]; then export PRO=1; else unset PRO; fi
    echo -n "$mode ${6:-raw} "; python tools/bench_conv.py fwd $1 $2 $3 $4 $5 3 20 2>&1 | tail -1
  done
done
