mkdir -p gpurun_out
timeout 400 python tools/exp_ab_option.py tc_ring_a 3 3,4 2>&1 | tail -6 | tee gpurun_out/r03j_ab.log
timeout 100 python - <<'PY' 2>&1 | grep -v Warn | tail -3
import sys; sys.argv=['x','tc_interleave=0']
PY
