#!/usr/bin/env python
"""Print selected metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` output (file argument)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__cycles_elapsed.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:70], r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    for k in KEYS + extra:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
