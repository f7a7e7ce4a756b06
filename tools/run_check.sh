mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r03r_bench.json 2> gpurun_out/r03r_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03r_bench_ref.json 2> gpurun_out/r03r_bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03r_bench.json').read().strip().splitlines()[-1])
print('steps', d['steps'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ratio', d['e2e']['value']/d['value'], d['clocks'])
print('config', d['config'])
print('roofline', {k: d['roofline'][k] for k in ('frac','frac_executed','avg_launch_ms','traffic','peak')})
print('cpu', d['cpu_baseline'])
c=d['config4']; print('config4', c['value'], c['wall_s'], c['per_rank'])
r=json.loads(open('gpurun_out/r03r_bench_ref.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['config']==d['config'], r.get('impl'))
PY
