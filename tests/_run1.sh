F="--set full --clock-control none --import-source on"
ncu $F -k regex:vgg_first_conv --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_vggfirst python tests/gpu_profile_conv.py vgg 8 > /dev/null 2>&1
ncu $F -k regex:s2conv --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_s2conv python tests/gpu_profile_conv.py s2 8 > /dev/null 2>&1
ncu $F -k regex:s2wgrad_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_s2wgrad python tests/gpu_profile_conv.py s2 8 > /dev/null 2>&1
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wgrad_accum" -s 2>&1 | tail -5
ls -la gpurun_out/*.ncu-rep
