cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -2
python tools/trace_probe.py c3 2 2>&1 | grep -E "total|crc"
