cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py > gpurun_out/r2_gputest_10.txt 2>&1; tail -5 gpurun_out/r2_gputest_10.txt
echo "== block search"; python tools/trace_probe.py c3 2 2>&1 | grep -E "it |total"
echo "== climb only"; NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap_noblk.so python tools/trace_probe.py c3 2 2>&1 | grep -E "it |total"
NW_PROFILE=3 python tools/trav_probe.py c3 2>&1 | grep "^iter"
