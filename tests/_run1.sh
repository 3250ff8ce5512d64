Q="--no-lpips-step --no-cpu-baseline --no-eager --no-disc-step --no-e2e"
for ch in 0 4096 2048; do
  echo "== FO_WG_CHAIN=$ch"
  FO_WG_CHAIN=$ch timeout 200 python tests/gpu_accum_bias.py 2>&1 | grep "one launch"
  FO_WG_CHAIN=$ch timeout 200 python bench.py $Q > gpurun_out/w_c32_$ch.json 2>> gpurun_out/w_err.log
  python -c "import json;d=json.load(open('gpurun_out/w_c32_$ch.json'));c=d['roofline_by_layer_class'];print('c32 chain$ch', d['ms_per_step'], d['clocks']['sm_mhz'], 'wgrad', d['kernels']['wgrad_igemm']['ms_per_step'], 'wgrad3d', c['wgrad3d_128x128']['ms_per_step'], 'w4x4', c['wgrad4x4s2_128x64']['ms_per_step'], 'w3x3_64', c['wgrad3x3_128x64']['ms_per_step'])"
done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
tail -5 gpurun_out/w_err.log
