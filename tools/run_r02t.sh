mkdir -p gpurun_out
timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 0 2>&1 | grep -v Warn | tail -16 | tee gpurun_out/r02t_trace.log
timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 8 2>&1 | grep -v Warn | grep -E "===|cta 0:|producer|epilogue" | tee -a gpurun_out/r02t_trace.log
timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 200 2>&1 | grep -v Warn | grep -E "===|cta 0:|producer|epilogue" | tee -a gpurun_out/r02t_trace.log
