mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench_n2.json 2> gpurun_out/r02z_bench_n2.err
echo "exit code $?"
wc -c gpurun_out/r02z_bench_n2.json gpurun_out/r02z_bench_n2.err
grep -v "no weights\|^\*\*\*\|OMP_NUM" gpurun_out/r02z_bench_n2.err | tail -20
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench_n2.json'))
print('n', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
c=d.get('config4',{}); print('config4', {k: c.get(k) for k in ('value','wall_s','lpt_imbalance','scaling')}); print(c.get('per_rank'))
PY
timeout 300 python tools/run_configs.py config5 2>&1 | tail -1 | tee gpurun_out/r02z_config5.json
