mkdir -p gpurun_out
cp mbexwn_vocoder_b200/libmbexwn_b200.so /tmp/lib_new.so
run() { cp $1 mbexwn_vocoder_b200/libmbexwn_b200.so; shift; timeout 200 python tools/exp_time_step.py "$@" 2>&1 | tail -1; }
for rep in 1 2; do
  run tools/attic/libmbexwn_b200_layer_r02l.so "r02l (4b9d940)"
  run tools/attic/libmbexwn_b200_layer_15226c4.so "15226c4 quad+hints" 
  run tools/attic/libmbexwn_b200_layer_15226c4.so "15226c4 hints off" tc_l2_hints=0
  run tools/attic/libmbexwn_b200_layer_bd47289.so "bd47289 interleave=0" tc_interleave=0
  run tools/attic/libmbexwn_b200_layer_1a996a7.so "1a996a7 spin build, interleave=0" tc_interleave=0
  run /tmp/lib_new.so "current interleave=0 discard=0" tc_interleave=0 tc_discard=0
  run /tmp/lib_new.so "current default"
done | tee gpurun_out/r03g_bisect.log
cp /tmp/lib_new.so mbexwn_vocoder_b200/libmbexwn_b200.so
