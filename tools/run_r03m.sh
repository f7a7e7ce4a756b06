mkdir -p gpurun_out
timeout 400 python tools/exp_ab_option.py tc_ring_a 3 2,3 2>&1 | tail -6 | tee gpurun_out/r03m_ab.log
timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 0 2>&1 | grep -v Warn | tail -16 | tee gpurun_out/r03m_trace.log
