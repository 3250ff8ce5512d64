timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "vgg or lpips or maxpool" 2>&1 | tail -4
Q="--no-lpips-step --no-cpu-baseline --no-eager --no-disc-step --no-e2e"
timeout 300 python bench.py --lpips 1 $Q > gpurun_out/f_c32l.json 2>> gpurun_out/f_err.log
python -c "import json;d=json.load(open('gpurun_out/f_c32l.json'));print('c32 lpips', d['ms_per_step'], d['clocks']['sm_mhz'], {k:(v['ms_per_step'],v['frac_of_hbm_peak']) for k,v in d['hbm_kernels'].items() if 'lpips' in k or 'pool' in k})"
tail -3 gpurun_out/f_err.log
