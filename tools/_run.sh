cd $GRAFT_REPO_ROOT
for o in 1.5 3 6 12; do echo -n "NW_LEAF_OCC=$o: "; NW_LEAF_OCC=$o python tools/trace_probe.py c3 2 2>&1 | grep -E "it 3|total" | sed -e 's/refit.*//' | tr '\n' ' '; echo; done
