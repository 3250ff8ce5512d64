#!/bin/bash
# vq_assign: A/B timing against another build (FACEOFF_B200_LIB) and ncu tensor-pipe utilisation at the sweep shapes
mkdir -p gpurun_out
if [ -f faceoff_b200/libfaceoff_b200_oldvq.so ]; then
  echo "== old vq.cu"; FACEOFF_B200_LIB=$PWD/faceoff_b200/libfaceoff_b200_oldvq.so python tests/gpu_profile_vq.py 32 64 512
fi
echo "== new"; python tests/gpu_profile_vq.py 32 64 512
for cfg in "64 512" "64 2048" "128 2048"; do
  set -- $cfg
  ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:vq_assign --launch-skip 3 --launch-count 1 python tests/gpu_profile_vq.py 32 $1 $2 2>&1 | grep -E "vq_assign_kernel|pipe_tensor|gpu__time|dram__bytes" | sed "s/^/[$1x$2] /"
done
ncu --set full --clock-control none --import-source on -k regex:vq_assign --launch-skip 3 --launch-count 1 -f -o gpurun_out/s2_vq64x512 python tests/gpu_profile_vq.py 32 64 512 > /dev/null 2>&1
