mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/q_tc_1x1 python tools/bench_conv.py fwd 16 128 128 224 128 1 1 > gpurun_out/q1.log 2>&1
$N -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/q_tc_dg python tools/bench_conv.py fwd 16 64 64 128 480 1 1 > gpurun_out/q2.log 2>&1
$N -k regex:conv_wgrad_tc -s 2 -c 1 -o gpurun_out/q_wg_1x1 python tools/bench_conv.py wgrad 16 64 64 256 128 1 1 > gpurun_out/q3.log 2>&1
$N -k regex:conv_wgrad_halo -s 2 -c 1 -o gpurun_out/q_wgh python tools/bench_conv.py wgrad 16 128 128 128 32 3 1 > gpurun_out/q4.log 2>&1
