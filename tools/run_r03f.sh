mkdir -p gpurun_out
cp mbexwn_vocoder_b200/libmbexwn_b200.so /tmp/lib_new.so
cp tools/attic/libmbexwn_b200_layer_r02l.so /tmp/lib_old.so
for rep in 1 2 3; do
  cp /tmp/lib_old.so mbexwn_vocoder_b200/libmbexwn_b200.so; timeout 200 python tools/exp_time_step.py "r02l layer kernel" 2>&1 | tail -1
  cp /tmp/lib_new.so mbexwn_vocoder_b200/libmbexwn_b200.so; timeout 200 python tools/exp_time_step.py "current (interleave + discard)" 2>&1 | tail -1
done | tee gpurun_out/r03f_ab_builds.log
cp /tmp/lib_new.so mbexwn_vocoder_b200/libmbexwn_b200.so
timeout 300 python tools/exp_ab_option.py tc_interleave 2 0,1 2>&1 | tail -4 | tee -a gpurun_out/r03f_ab_builds.log
timeout 300 python tools/exp_ab_option.py tc_fused 2 0,1 2>&1 | tail -4 | tee -a gpurun_out/r03f_ab_builds.log
