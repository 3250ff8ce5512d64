#!/bin/bash
# Round-2 (second half) profiling evidence, run under gpurun; outputs in gpurun_out/, summarised into profiles/ by
# profiles/summarize.py.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline --no-eager --no-disc-step"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_cfg1.csv python bench.py --steps 2 --warmup 1 $B --no-lpips-step > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_cfg2.csv python bench.py --steps 2 --warmup 1 $B --no-lpips-step --lpips 1 > /dev/null 2>&1
F="--set full --clock-control none --import-source on"
ncu $F -k regex:conv_igemm --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2b_conv3d python tests/gpu_profile_conv.py conv3d 32 > /dev/null 2>&1
ncu $F -k regex:wgrad_igemm --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2b_wgrad3d python tests/gpu_profile_conv.py wgrad3d 32 > /dev/null 2>&1
ncu $F -k regex:wgrad_igemm --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2b_wgrad4x4s2 python tests/gpu_profile_conv.py wgrad_down 32 > /dev/null 2>&1
ncu $F -k regex:wgrad_igemm --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2b_wgrad1x1 python tests/gpu_profile_conv.py wgrad_small 32 > /dev/null 2>&1
ncu $F -k regex:vq_assign --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2b_vq64x512 python tests/gpu_profile_vq.py 32 64 512 > /dev/null 2>&1
ncu $F -k regex:conv_igemm --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2b_vgg64 python tests/gpu_profile_conv.py vgg 8 > /dev/null 2>&1
for cfg in "64 512" "64 1024" "64 2048" "128 512" "128 1024" "128 2048"; do
  set -- $cfg
  ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum --clock-control none -k regex:vq_assign --launch-skip 3 --launch-count 1 python tests/gpu_profile_vq.py 32 $1 $2 2>&1 | grep -E "pipe_tensor|gpu__time" | sed "s/^/[$1x$2] /"
done > gpurun_out/r2b_vq_tensor_pct.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2b_hbm_kernels.csv python tests/gpu_profile_hbm.py 8 > /dev/null 2>&1
python tests/gpu_profile_hbm.py 32 --time > gpurun_out/r2b_hbm_time.txt 2>&1
python tests/gpu_profile_vq.py 32 > gpurun_out/r2b_vq_sweep.txt 2>&1
python tests/gpu_profile_conv.py all 32 > gpurun_out/r2b_conv_isolated.txt 2>&1
python tests/gpu_profile_disc.py all > gpurun_out/r2b_disc.txt 2>&1
ls -la gpurun_out | tail -20
