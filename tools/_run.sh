cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py > gpurun_out/r2_gputest_4.txt 2>&1; tail -4 gpurun_out/r2_gputest_4.txt
python tools/trace_probe.py c3 2 2>&1 | tail -13
for v in "" _cv4 _cv5; do NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap$v.so python tools/kernel_probe.py curvature 2>&1 | tail -1 | tr '\n' ' '; echo " [$v]"; done
python tools/kernel_probe.py apply_A apply_AH > gpurun_out/plain_kp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_apply_A" -c 6 -o gpurun_out/r2_prof_adjoint python tools/kernel_probe.py apply_A apply_AH > gpurun_out/ncu_kp.log 2>&1; tail -2 gpurun_out/ncu_kp.log
