timeout 300 python tools/exp_ab_option.py tc_ring_a 2 3,2 2>&1 | tail -4
