cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -2
NW_TRACE_BUILD=1 python tools/trace_probe.py c3 2 2>&1 | grep -E "normals|it 3|total|crc" | tail -6
python tools/trace_probe.py c5 2 2>&1 | grep -E "block|total|crc"
