# round-2 evidence: launch list of one training step + full captures of the kernels added / changed this round (run under gpurun)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python tools/one_step.py > gpurun_out/r2_0.log 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:conv_halo_tma -s 3 -c 1 -o gpurun_out/r2_halo_tma_raw64 python tools/bench_conv.py fwd 16 256 256 64 64 3 1 > gpurun_out/r2_1.log 2>&1
PRO=1 $N -k regex:conv_halo_tma -s 3 -c 1 -o gpurun_out/r2_halo_tma_pro32 python tools/bench_conv.py fwd 16 64 64 128 32 3 1 > gpurun_out/r2_2.log 2>&1
$N -k regex:conv_halo_tma -s 3 -c 1 -o gpurun_out/r2_halo_tma_n128 python tools/bench_conv.py fwd 16 128 128 32 128 3 1 > gpurun_out/r2_3.log 2>&1
$N --profile-from-start off -k regex:conv_pw_t -s 81 -c 1 -o gpurun_out/r2_pwt_bnbwd python tools/one_step.py > gpurun_out/r2_4.log 2>&1
$N --profile-from-start off -k regex:"bn_fixup|optimizer_step|edge_gt|argmax" -c 2 -o gpurun_out/r2_small python tools/one_step.py > gpurun_out/r2_5.log 2>&1
