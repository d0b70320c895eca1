mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
NOSTAT=1 $N -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r_k64 python tools/bench_conv.py fwd 16 128 128 64 128 1 1 > gpurun_out/r1.log 2>&1
NOSTAT=1 $N -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r_n32 python tools/bench_conv.py fwd 16 128 128 224 32 1 1 > gpurun_out/r2.log 2>&1
