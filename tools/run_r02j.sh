mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02j_pytest_gpu.log
cat gpurun_out/r02j_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
tail -3 gpurun_out/r02j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'cabi', d['e2e']['cabi']['value'], 'serial', d['e2e']['serial']['value'])
print('config4', {k:v for k,v in d.get('config4',{}).items() if k not in ('shard_audio_s',)})
print('cpu', d.get('cpu_baseline'))
PY
