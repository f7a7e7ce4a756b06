mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_full_size.py -q -x 2>&1 | tail -3
timeout 400 python tools/exp_ab_option.py tc_discard 2 0,1 2>&1 | tail -4 | tee gpurun_out/r03c_ab_discard.log
M=gpu__time_duration.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for c in "tc_discard=0" "tc_discard=1" "tc_discard=1 tc_interleave=1"; do
timeout 300 ncu -k regex:wn_layer -s 2 -c 2 --metrics $M --clock-control none --csv --log-file "gpurun_out/r03c_$(echo $c | tr ' =' '__').csv" python tools/exp_one_forward.py $c 2>&1 | grep -v Warn | tail -1
done
ls gpurun_out | grep r03c
