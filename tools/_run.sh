cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_2.txt 2>&1; tail -5 gpurun_out/r2_gputest_2.txt
python tools/trace_probe.py c3 4 > gpurun_out/r2_trace_c3.txt 2>&1; tail -30 gpurun_out/r2_trace_c3.txt
