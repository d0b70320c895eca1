#!/bin/bash
# round-2 final evidence (run under gpurun, 1 GPU): launch lists of one fp32-class and one bf16 training step, full captures
# of the kernels changed by the bf16 path, and the bench lines the docs quote.  Every command has its own timeout.
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f2_launches_f32.csv python tools/one_step.py > gpurun_out/f2_0.log 2>&1
SAUNET_PRECISION=bf16 timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f2_launches_bf16.csv python tools/one_step.py > gpurun_out/f2_1.log 2>&1
SAUNET_PRECISION=bf16 timeout 120 $N -k regex:conv_halo_tma -s 3 -c 1 -o gpurun_out/f2_halo_tma_bf16_raw64 python tools/bench_conv.py fwd 16 256 256 64 64 3 1 > gpurun_out/f2_2.log 2>&1
SAUNET_PRECISION=bf16 timeout 120 $N -k regex:conv_halo_tma -s 3 -c 1 -o gpurun_out/f2_halo_tma_bf16_n128 python tools/bench_conv.py fwd 16 64 64 256 128 3 1 > gpurun_out/f2_3.log 2>&1
timeout 300 python bench.py > gpurun_out/f2_bench_1gpu_f32.json 2> gpurun_out/f2_4.log
timeout 300 python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/f2_bench_1gpu_bf16_b32.json 2> gpurun_out/f2_5.log
timeout 300 python bench.py --dtype bf16 --workload train_loop --no-cpu-baseline > gpurun_out/f2_loop_1gpu_bf16_b32.json 2> gpurun_out/f2_6.log
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f2_ref.json 2> gpurun_out/f2_7.log
tail -c 300 gpurun_out/f2_[0-7].log
cut -c1-300 gpurun_out/f2_bench_1gpu_f32.json gpurun_out/f2_bench_1gpu_bf16_b32.json gpurun_out/f2_loop_1gpu_bf16_b32.json gpurun_out/f2_ref.json
