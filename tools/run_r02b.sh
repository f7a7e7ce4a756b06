mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "fused or parity" 2>&1 | tail -5 > gpurun_out/r02b_pytest_tc.log
cat gpurun_out/r02b_pytest_tc.log
for d in 0 1 2 3 4 8 12 7; do timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 $d 2>&1 | grep -v Warn | tail -8; done | tee gpurun_out/r02b_trace_debug.log
