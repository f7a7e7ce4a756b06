mkdir -p gpurun_out
timeout 600 python tools/exp_ab_option.py tc_interleave 5 0,1 2>&1 | tail -10 | tee gpurun_out/r03d_ab_interleave.log
