mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r03p_bench_n$N.json 2> gpurun_out/r03p_bench_n$N.err
echo "exit $?"
python - <<PY
import json
txt=open('gpurun_out/r03p_bench_n$N.json').read()
lines=[l for l in txt.splitlines() if l.startswith('{')]
print('stdout lines:', len(txt.splitlines()))
d=json.loads(lines[-1])
print('n', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])
c=d.get('config4',{}); print('config4', {k: c.get(k) for k in ('value','wall_s','lpt_imbalance','scaling','host_threads_per_rank')}); print(c.get('per_rank'))
PY
