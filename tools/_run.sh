cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py > gpurun_out/r2_gputest_7.txt 2>&1; tail -12 gpurun_out/r2_gputest_7.txt
