cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py > gpurun_out/t.txt 2>&1; tail -2 gpurun_out/t.txt
python tools/trace_probe.py c3 3 > gpurun_out/tp.txt 2>&1; grep -E "it 0|total|crc" gpurun_out/tp.txt
