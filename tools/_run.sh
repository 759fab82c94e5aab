cd $GRAFT_REPO_ROOT
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_edge_cases.py -x -q -k "zero_points" > gpurun_out/r2_sanitizer_zero.txt 2>&1; grep -v "^$" gpurun_out/r2_sanitizer_zero.txt | head -60
python tools/moves_probe.py c3 80 > gpurun_out/r2_moves_c3.txt 2>&1; cat gpurun_out/r2_moves_c3.txt | tail -85
