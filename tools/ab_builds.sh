# Same-box A/B of library builds (alternating processes): the r02l layer kernel kept under tools/attic/ (build it from commit 4b9d940 if absent),
# the current library with several option sets, and the two-launch path.  Usage under gpurun: bash tools/ab_builds.sh
mkdir -p gpurun_out
cp mbexwn_vocoder_b200/libmbexwn_b200.so /tmp/lib_new.so
run() { cp $1 mbexwn_vocoder_b200/libmbexwn_b200.so; shift; timeout 200 python tools/exp_time_step.py "$@" 2>&1 | tail -1; }
for rep in 1 2; do
  [ -f tools/attic/libmbexwn_b200_layer_r02l.so ] && run tools/attic/libmbexwn_b200_layer_r02l.so "r02l (4b9d940)"
  run /tmp/lib_new.so "current interleave=0 discard=0" tc_interleave=0 tc_discard=0
  run /tmp/lib_new.so "current interleave=0 discard=1" tc_interleave=0 tc_discard=1
  run /tmp/lib_new.so "current default (interleave=1 discard=1)"
  run /tmp/lib_new.so "unfused" tc_fused=0
done | tee gpurun_out/r03h_bisect.log
cp /tmp/lib_new.so mbexwn_vocoder_b200/libmbexwn_b200.so
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -3
