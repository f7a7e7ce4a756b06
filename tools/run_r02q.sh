mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -5
timeout 400 python tools/exp_ab_option.py tc_l2_hints 2 0,1,2,3 2>&1 | tail -8 | tee gpurun_out/r02q_ab_hints.log
M=gpu__time_duration.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for c in 0 1 3; do
timeout 300 ncu -k regex:wn_layer -s 2 -c 2 --metrics $M --clock-control none --csv --log-file gpurun_out/r02q_hints$c.csv python tools/exp_one_forward.py tc_l2_hints=$c 2>&1 | grep -v Warn | tail -1
done
