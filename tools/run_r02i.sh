timeout 300 python tools/exp_fused_diag.py f16f8 400,380,17,400,211,400,1,400,399,400,2,400,400,333 2>&1 | grep -v Warn | tail -30
timeout 300 python tools/exp_fused_diag.py f16f8 400,150,1,400,37,400,400,260 2>&1 | grep -v Warn | tail -12
