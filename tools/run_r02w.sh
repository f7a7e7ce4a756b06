mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
timeout 400 python tools/exp_ab_option.py tc_interleave 2 0,1 2>&1 | tail -4 | tee gpurun_out/r02w_ab_interleave.log
timeout 400 python tools/exp_ab_option.py tc_ring_a 1 2,3 2>&1 | tail -2 | tee gpurun_out/r02w_ab_ring.log
M=gpu__time_duration.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for c in 0 1; do
timeout 300 ncu -k regex:wn_layer -s 2 -c 2 --metrics $M --clock-control none --csv --log-file gpurun_out/r02w_il$c.csv python tools/exp_one_forward.py tc_interleave=$c 2>&1 | grep -v Warn | tail -1
done
