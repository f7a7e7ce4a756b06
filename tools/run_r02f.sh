mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -k "fused" 2>&1 | tail -3
timeout 300 python tools/exp_ab_option.py tc_fused 2 0,1 2>&1 | tail -4 | tee gpurun_out/r02f_ab_fused.log
for d in 0 4; do timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 $d 2>&1 | grep -v Warn | tail -14; done | tee gpurun_out/r02f_trace.log
