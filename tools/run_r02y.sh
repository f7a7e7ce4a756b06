mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 600 python tools/run_configs.py config5 2>&1 | tail -1 | tee gpurun_out/r02y_config5.json
timeout 600 python tools/run_configs.py config5 --max-batch-frames 24576 2>&1 | tail -1 | tee gpurun_out/r02y_config5_24k.json
