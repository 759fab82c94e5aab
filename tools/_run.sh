cd $GRAFT_REPO_ROOT
NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap_bounds.so timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_bounds.txt 2>&1; tail -3 gpurun_out/r2_gputest_bounds.txt
NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap_bounds.so python tools/trace_probe.py c3 1 2>&1 | grep -E "total|crc"
python tools/trace_probe.py c3 1 2>&1 | grep -E "total|crc"
