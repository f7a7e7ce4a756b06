mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02l_pytest_gpu.log
cat gpurun_out/r02l_pytest_gpu.log
timeout 600 python tools/run_configs.py config5 2>&1 | tail -1 | tee gpurun_out/r02l_config5.json
timeout 600 python bench.py --workload config1 --steps 200 --warmup 5 --no-cpu-baseline --no-config4 2>gpurun_out/r02l_c1.err | python -c "import json,sys; d=json.load(sys.stdin); print('config1 ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['serial']['ms_per_step'])"
bash tools/profile_round.sh r02l
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
c=d.get('config4',{}); print('config4', c.get('value'), c.get('wall_s'), c.get('per_rank'))
PY
