# per-kernel device times of the WaveNet GEMMs under the tc_debug timing experiments (ncu launch lists)
for d in "$@"; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wn_gemm -s 48 -c 16 --csv --log-file gpurun_out/exp_d$d.csv python bench.py --precision f16f8 --steps 1 --warmup 3 --no-cpu-baseline --tc-debug $d > /dev/null 2>&1
  echo "debug $d"; python tools/ncu_launch_summary.py gpurun_out/exp_d$d.csv | grep "EPI=[12]"
done
