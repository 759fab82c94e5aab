cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py > gpurun_out/r2_gputest_8.txt 2>&1; tail -3 gpurun_out/r2_gputest_8.txt
for wl in c1 c2; do python tools/graph_probe.py $wl 2>&1 | tail -2; NW_NO_GRAPH=1 python tools/graph_probe.py $wl 2>&1 | tail -2; done
