mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py 2>&1 | grep -v "Warn\|warning::" | tail -6 | tee gpurun_out/r03q_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py blocks 2>&1 | grep -v "Warn\|warning::" | tail -4 | tee gpurun_out/r03q_memcheck_blocks.log
timeout 900 compute-sanitizer --tool initcheck --print-limit 5 python tools/sanitize_small.py 2>&1 | grep -v "Warn\|warning::" | tail -4 | tee gpurun_out/r03q_initcheck.log
