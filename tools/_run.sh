cd $GRAFT_REPO_ROOT
python tools/shard_probe3.py 2>&1 | grep -E "cubes|Error"
