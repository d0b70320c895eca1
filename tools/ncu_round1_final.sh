# round-1 final evidence: launch list of one training step + full captures of the dominant kernels (run under gpurun)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f_launches.csv python tools/one_step.py > gpurun_out/f0.log 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:conv_wgrad_halo -s 2 -c 1 -o gpurun_out/f_wgrad_halo python tools/bench_conv.py wgrad 16 128 128 128 32 3 1 > gpurun_out/f1.log 2>&1
$N -k regex:conv_wgrad_pw -s 2 -c 1 -o gpurun_out/f_wgrad_pw python tools/bench_conv.py wgrad 16 64 64 480 128 1 1 > gpurun_out/f2.log 2>&1
$N -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/f_tc_1x1 python tools/bench_conv.py fwd 16 64 64 480 128 1 1 > gpurun_out/f3.log 2>&1
$N -k regex:conv_halo -s 3 -c 1 -o gpurun_out/f_halo_n32 python tools/bench_conv.py fwd 16 128 128 128 32 3 1 > gpurun_out/f4.log 2>&1
$N -k regex:conv_halo -s 3 -c 1 -o gpurun_out/f_halo_n64 python tools/bench_conv.py fwd 16 256 256 64 64 3 1 > gpurun_out/f5.log 2>&1
$N --profile-from-start off -k regex:"bn_bwd_apply4|bn_bwd_reduce4" -s 40 -c 2 -o gpurun_out/f_bn python tools/one_step.py > gpurun_out/f6.log 2>&1
