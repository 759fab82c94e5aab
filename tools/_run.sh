cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -2
python tools/trace_probe.py c3 2 2>&1 | grep -E "it |total|crc"
python tools/vorder_probe.py c3 2>&1 | grep -E "vertices|Error"
