mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python tools/run_configs.py config5 2>&1 | tail -1 | tee gpurun_out/r03b_config5.json
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err
timeout 900 python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/r03b_bench_config3.json 2> gpurun_out/r03b_bench_config3.err
python - <<'PY'
import json
for f in ('bench','bench_config3'):
    d=json.load(open(f'gpurun_out/r03b_{f}.json'))
    print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ratio', d['e2e']['value']/d['value'], 'cabi', d['e2e']['cabi']['value'], 'traffic', d['roofline']['traffic'], d['clocks']['sm_mhz'])
PY
