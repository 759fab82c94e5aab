cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py > gpurun_out/r2_gputest_9.txt 2>&1; tail -3 gpurun_out/r2_gputest_9.txt
NW_TRACE_BUILD=1 python tools/trace_probe.py c3 3 2>&1 | grep -E "feet \+ uploads" | tail -3
python tools/e2e_profile.py 2>&1 | tail -8
python tools/curv_probe.py c3 2>&1 | tail -4
