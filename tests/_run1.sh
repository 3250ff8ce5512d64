timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "vgg or lpips or maxpool" 2>&1 | tail -4
timeout 300 python tests/gpu_profile_conv.py vgg 32 2>&1 | grep "first conv"
