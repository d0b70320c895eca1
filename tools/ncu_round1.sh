mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:conv_wgrad_tc -s 2 -c 1 -o gpurun_out/p_wg_d1 python tools/bench_conv.py wgrad 16 128 128 128 32 3 1 > gpurun_out/p1.log 2>&1
$N -k regex:conv_wgrad_tc -s 2 -c 1 -o gpurun_out/p_wg_r1 python tools/bench_conv.py wgrad 16 256 256 64 64 3 1 > gpurun_out/p2.log 2>&1
$N -k regex:conv_halo -s 2 -c 1 -o gpurun_out/p_halo_d1 python tools/bench_conv.py fwd 16 128 128 128 32 3 1 > gpurun_out/p3.log 2>&1
$N -k regex:conv_halo -s 2 -c 1 -o gpurun_out/p_halo_r1 python tools/bench_conv.py fwd 16 256 256 64 64 3 1 > gpurun_out/p4.log 2>&1
$N -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/p_tc_1x1 python tools/bench_conv.py fwd 16 128 128 224 128 1 1 > gpurun_out/p5.log 2>&1
$N --profile-from-start off -k regex:"bn_bwd_apply|chan_reduce" -s 20 -c 4 -o gpurun_out/p_bn python tools/one_step.py > gpurun_out/p6.log 2>&1
for g in "wgrad 16 128 128 128 32 3" "wgrad 16 256 256 64 64 3" "wgrad 16 64 64 128 32 3" "wgrad 16 32 32 128 32 3" "fwd 16 128 128 128 32 3" "fwd 16 256 256 64 64 3" "fwd 16 128 128 224 128 1" "fwd 16 32 32 512 128 1" "fwd 16 128 128 32 128 3"; do python tools/bench_conv.py $g 20; done > gpurun_out/micro.log 2>&1
cat gpurun_out/micro.log
