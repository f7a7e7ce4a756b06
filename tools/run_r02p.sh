mkdir -p gpurun_out
M=gpu__time_duration.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max
for c in 2 4; do
timeout 300 ncu -k regex:wn_layer -s 2 -c 2 --metrics $M --clock-control none --csv --log-file gpurun_out/r02p_cluster$c.csv python tools/exp_one_forward.py tc_cluster=$c 2>&1 | grep -v Warn | tail -2
done
