cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests/test_gpu_scale.py -x -q -s > gpurun_out/r2_gputest_scale.txt 2>&1; tail -30 gpurun_out/r2_gputest_scale.txt
