mkdir -p gpurun_out
for p in f16f8 bf16x3; do timeout 300 python tools/exp_fused_diag.py $p 2>&1 | grep -v Warn | tail -8; done | tee gpurun_out/r02d_diag.log
for d in 0 1 2 16; do timeout 200 python tools/exp_trace_layer.py 3 "" f16f8 config2 $d 2>&1 | grep -v Warn | grep -E "===|cta 0:|producer|epilogue" | head -4; done | tee gpurun_out/r02d_trace.log
