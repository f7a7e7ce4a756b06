#!/bin/bash
# Profiles of one bench step under gpurun (1 GPU): launch list of every kernel + `--set full` of the fused WaveNet layer kernel
# (one mid-stack launch), of the sub-net tap-GEMMs and of the memory-bound kernels.
# Usage: tools/profile_round.sh <tag> [bench args...]   -> gpurun_out/<tag>_*.{csv,ncu-rep}
tag=$1; shift
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config4"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_all.csv \
    $B "$@" > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:wn_layer_kernel' -s 27 -c 2 -o gpurun_out/${tag}_wn_layer \
    $B "$@" > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k 'regex:stft_filter2048_kernel|post_pqmf_kernel|pulse_kernel|phase_chunk_kernel|ola_kernel|start_pack_kernel' -s 18 -c 6 -o gpurun_out/${tag}_membound \
    $B "$@" > /dev/null 2>&1
for f in wn_layer membound; do
  ncu -i gpurun_out/${tag}_${f}.ncu-rep --page raw --csv > gpurun_out/${tag}_${f}_full_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep ${tag}
