#!/bin/bash
# ncu --set full captures of the VGG block-1 launches (run under gpurun)
mkdir -p gpurun_out
python tests/gpu_profile_conv.py vgg 8
ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 12 --launch-count 1 -f -o gpurun_out/s2_vgg64 python tests/gpu_profile_conv.py vgg 8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 25 --launch-count 1 -f -o gpurun_out/s2_vggfirst python tests/gpu_profile_conv.py vgg 8 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
