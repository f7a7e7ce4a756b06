#!/bin/bash
# Profiles of one bench step under gpurun (1 GPU): launch list of every kernel + `--set full` of one gate and one res/skip
# tap-GEMM launch.  Usage: tools/profile_round.sh <tag> [bench args...]   -> gpurun_out/<tag>_*.{csv,ncu-rep}
tag=$1; shift
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_all.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:wn_gemm_kernel<\(int\)[12]' -s 52 -c 2 -o gpurun_out/${tag}_wn_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on \
    -k 'regex:stft_filter2048_kernel|post_pqmf_kernel|pulse_kernel|phase_chunk_kernel' -s 12 -c 4 -o gpurun_out/${tag}_membound \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2>&1
ls -la gpurun_out | grep ${tag}
