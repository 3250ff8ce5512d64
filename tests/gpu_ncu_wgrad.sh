#!/bin/bash
# ncu --set full captures of the wgrad kernel at the shapes that are below roofline (run under gpurun)
mkdir -p gpurun_out
python tests/gpu_profile_conv.py wgrad_small 32
python tests/gpu_profile_conv.py wgrad_down 32
ncu --set full --clock-control none --import-source on -k regex:wgrad_igemm --launch-skip 13 --launch-count 1 -f -o gpurun_out/s2_wg_down python tests/gpu_profile_conv.py wgrad_down 32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_igemm --launch-skip 13 --launch-count 1 -f -o gpurun_out/s2_wg_1x1 python tests/gpu_profile_conv.py wgrad_small 32 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
