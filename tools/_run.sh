cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_golden.py -x -q -k "block_driver" 2>&1 | tail -15
