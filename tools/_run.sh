cd $GRAFT_REPO_ROOT
python tools/deep_probe.py 2>&1 | grep -E "slab|without|quantiles|Error"
