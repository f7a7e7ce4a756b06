mkdir -p gpurun_out
M=gpu__time_duration.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for d in 0 1 2 64 128 192 194 8; do
timeout 300 ncu -k regex:wn_layer -s 2 -c 1 --metrics $M --clock-control none --csv --log-file gpurun_out/r02r_dbg$d.csv python tools/exp_one_forward.py tc_trace=3 tc_debug=$d 2>&1 | grep -v Warn | tail -1
done
