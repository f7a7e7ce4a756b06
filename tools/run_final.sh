mkdir -p gpurun_out
T=r03n
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${T}_pytest_gpu.log
cat gpurun_out/${T}_pytest_gpu.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/profile_round.sh ${T} > /dev/null
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wn_layer_kernel' -s 27 -c 1 \
    -o gpurun_out/${T}_wn_layer_config3 python bench.py --workload config3 --steps 1 --warmup 3 --no-cpu-baseline --no-config4 > /dev/null 2>&1
ncu -i gpurun_out/${T}_wn_layer_config3.ncu-rep --page raw --csv > gpurun_out/${T}_wn_layer_config3_full_raw.csv 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --fused 0 --no-cpu-baseline --no-config4 > gpurun_out/${T}_bench_unfused.json 2> gpurun_out/${T}_bench_unfused.err
timeout 600 python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/${T}_bench_config3.json 2> gpurun_out/${T}_bench_config3.err
timeout 600 python bench.py --workload config1 --steps 200 --warmup 5 --no-cpu-baseline --no-config4 > gpurun_out/${T}_bench_config1.json 2>gpurun_out/${T}_c1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python - <<'PY'
import json
for f in ('bench','bench_unfused','bench_config3','bench_config1','bench_ref'):
    try:
        d=json.load(open(f'gpurun_out/r03n_{f}.json'))
        print(f, 'ms/step', d.get('ms_per_step'), 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'), 'traffic', d.get('roofline',{}).get('traffic'), d.get('clocks'))
        c=d.get('config4'); 
        if c: print('  config4', c.get('value'), c.get('wall_s'), c.get('per_rank'))
    except Exception as e:
        print(f, 'failed', e)
PY
rm -f gpurun_out/${T}_*.ncu-rep
ls gpurun_out | grep ${T}
