mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_multi.py -q -x 2>&1 | tail -15
timeout 300 python tools/exp_ab_option.py tc_ring_a 2 2,3,4 2>&1 | tail -6 | tee gpurun_out/r02u_ab_ring.log
timeout 300 python tools/exp_ab_option.py tc_fused 1 0,1 2>&1 | tail -2 | tee gpurun_out/r02u_ab_fused.log
timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 0 2>&1 | grep -v Warn | tail -16 | tee gpurun_out/r02u_trace.log
