mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_full_size.py -q 2>&1 | tail -4
timeout 600 python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/r03t_bench_config3.json 2> gpurun_out/r03t_bench_config3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03t_bench_config3.json').read().strip().splitlines()[-1])
print('config3 ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'])
PY
