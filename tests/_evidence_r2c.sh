#!/bin/bash
# Round-2 (last third) evidence, run under gpurun from the repo root; outputs in gpurun_out/, summarised into profiles/r2c_*
# by profiles/summarize.py.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_tests.log 2>&1; tail -2 gpurun_out/r2c_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
B="--no-e2e --no-cpu-baseline --no-eager --no-disc-step"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_cfg1.csv python bench.py --steps 2 --warmup 1 $B --no-lpips-step > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_cfg2.csv python bench.py --steps 2 --warmup 1 $B --no-lpips-step --lpips 1 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_cfg2_4clips.csv python bench.py --clips 4 --steps 2 --warmup 1 $B --no-lpips-step --lpips 1 > /dev/null 2>&1
F="--set full --clock-control none --import-source on"
timeout 300 ncu $F -k regex:wgrad_igemm --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_wgrad3d python tests/gpu_profile_conv.py wgrad3d 32 > /dev/null 2>&1
timeout 300 ncu $F -k regex:vgg_first_conv --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_vggfirst python tests/gpu_profile_conv.py vgg 8 > /dev/null 2>&1
timeout 300 ncu $F -k regex:s2conv --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_s2conv python tests/gpu_profile_conv.py s2 8 > /dev/null 2>&1
timeout 300 ncu $F -k regex:s2wgrad_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_s2wgrad python tests/gpu_profile_conv.py s2 8 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"vgg_first_dgrad|tap_bwd_pool" --csv --log-file gpurun_out/r2c_new_hbm.csv python tests/gpu_profile_new.py > /dev/null 2>&1
timeout 200 python tests/gpu_accum_bias.py > gpurun_out/r2c_accum_bias.txt 2>&1
