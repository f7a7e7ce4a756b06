mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02h_pytest_gpu.log
cat gpurun_out/r02h_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
cat gpurun_out/r02h_bench.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --fused 0 > gpurun_out/r02h_bench_unfused.json 2>> gpurun_out/r02h_bench.err
cat gpurun_out/r02h_bench_unfused.json | python -c "import json,sys; d=json.load(sys.stdin); print('unfused', d['ms_per_step'], d['value'])"
