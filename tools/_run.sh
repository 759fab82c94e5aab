cd $GRAFT_REPO_ROOT
for d in 64 32 8 4 2; do echo "== far_div $d"; NW_COLD_FAR_DIV=$d python tools/trace_probe.py c3 1 2>&1 | grep -E "it 0"; done
echo "== no split, twice"; NW_NO_COLD_SPLIT=1 python tools/trace_probe.py c3 1 2>&1 | grep -E "it 0"; NW_NO_COLD_SPLIT=1 python tools/trace_probe.py c3 1 2>&1 | grep -E "it 0"
