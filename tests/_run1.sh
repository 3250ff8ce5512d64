Q="--no-lpips-step --no-cpu-baseline --no-eager --no-disc-step --no-e2e"
for ss in 1 0; do for pdl in 1 0; do
  FO_BENCH_SIDE_STREAM=$ss FO_PDL=$pdl timeout 200 python bench.py --clips 4 $Q --steps 20 > gpurun_out/s_c4_$ss$pdl.json 2>> gpurun_out/s_err.log
  python -c "import json;d=json.load(open('gpurun_out/s_c4_$ss$pdl.json'));print('c4 side$ss pdl$pdl', d['ms_per_step'], d['host_enqueue_ms_per_step'], d['clocks']['sm_mhz'], d['kernels']['wgrad_igemm']['ms_per_step'], d['roofline_by_layer_class']['wgrad3d_128x128']['ms_per_step'])"
done; done
tail -5 gpurun_out/s_err.log
