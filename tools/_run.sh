cd $GRAFT_REPO_ROOT
echo "== build trace"; NW_TRACE_BUILD=1 python tools/trace_probe.py c3 2 2>&1 | grep -E "nw build|block" | tail -24
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_c3_v2.csv python bench.py --steps 5 --warmup 5 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log | cut -c1-300
ncu --set full --import-source on --clock-control none -k regex:"k_adjoint|k_sweep2|k_apply_A" -c 6 -o gpurun_out/r2_prof_ops2 -f python tools/kernel_probe.py apply_A apply_AH adjoint sweep2 > gpurun_out/ncu_ops2.log 2>&1; tail -3 gpurun_out/ncu_ops2.log
