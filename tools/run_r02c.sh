mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -k "fused or parity" 2>&1 | tail -15 > gpurun_out/r02c_pytest_tc.log
cat gpurun_out/r02c_pytest_tc.log
timeout 300 python tools/exp_ab_option.py tc_fused 2 0,1 2>&1 | tail -4 | tee gpurun_out/r02c_ab_fused.log
timeout 300 python tools/exp_ab_option.py tc_slab 1 0,1 2>&1 | tail -2 | tee -a gpurun_out/r02c_ab_fused.log
for d in 0 3 4; do timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 $d 2>&1 | grep -v Warn | tail -22; done | tee gpurun_out/r02c_trace.log
