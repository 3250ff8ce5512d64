timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "lpips" 2>&1 | tail -3
timeout 200 python tests/gpu_profile_new.py 32 2>&1 | grep tap_bwd_pool
