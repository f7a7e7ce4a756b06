mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -3
timeout 400 python tools/exp_ab_option.py tc_spin 3 0,1 2>&1 | tail -6 | tee gpurun_out/r03a_ab_spin.log
