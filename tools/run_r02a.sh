set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -25 > gpurun_out/r02a_pytest_tc.log
cat gpurun_out/r02a_pytest_tc.log
timeout 300 python tools/exp_ab_option.py tc_fused 2 0,1 2>&1 | tail -8 | tee gpurun_out/r02a_ab_fused.log
timeout 300 python tools/exp_trace_layer.py 3 gpurun_out/r02a_trace.npy 2>&1 | tail -60 | tee gpurun_out/r02a_trace.log
