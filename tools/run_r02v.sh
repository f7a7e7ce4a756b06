mkdir -p gpurun_out
T=r02v
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest_gpu.log
cat gpurun_out/${T}_pytest_gpu.log | tail -8
timeout 600 python tools/run_configs.py config5 2>&1 | tail -2 | tee gpurun_out/${T}_config5.json
timeout 600 python bench.py --workload config1 --steps 200 --warmup 5 --no-cpu-baseline --no-config4 > gpurun_out/${T}_bench_config1.json 2>gpurun_out/${T}_c1.err
bash tools/profile_round.sh ${T} > /dev/null
# config 3 (bf16, C = 340): launch list + one full capture of the layer kernel + bench line
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wn_layer_kernel' -s 27 -c 1 \
    -o gpurun_out/${T}_wn_layer_config3 python bench.py --workload config3 --steps 1 --warmup 3 --no-cpu-baseline --no-config4 > /dev/null 2>&1
ncu -i gpurun_out/${T}_wn_layer_config3.ncu-rep --page raw --csv > gpurun_out/${T}_wn_layer_config3_full_raw.csv 2>/dev/null
timeout 600 python bench.py --workload config3 --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/${T}_bench_config3.json 2> gpurun_out/${T}_bench_config3.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --fused 0 --no-cpu-baseline --no-config4 > gpurun_out/${T}_bench_unfused.json 2> gpurun_out/${T}_bench_unfused.err
python - <<'PY'
import json
for f in ('bench','bench_unfused','bench_config3','bench_config1'):
    try:
        d=json.load(open(f'gpurun_out/r02v_{f}.json'))
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])
        c=d.get('config4'); 
        if c: print('  config4', c.get('value'), c.get('wall_s'), c.get('per_rank'))
    except Exception as e:
        print(f, 'failed', e)
PY
ls gpurun_out | grep r02v
