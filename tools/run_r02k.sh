mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02k_pytest_gpu.log
cat gpurun_out/r02k_pytest_gpu.log
timeout 600 python tools/run_configs.py config5 2>&1 | tail -2 | tee gpurun_out/r02k_config5.json
timeout 600 python tools/run_configs.py config5 2>&1 | tail -1 | tee -a gpurun_out/r02k_config5.json
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
c=d.get('config4',{}); print('config4', c.get('value'), c.get('wall_s'), c.get('per_rank'))
PY
