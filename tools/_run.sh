cd $GRAFT_REPO_ROOT
for n in 0 4 16 64 152; do echo -n "NW_COLD_SINGLE=$n: "; NW_COLD_SINGLE=$n python tools/trace_probe.py c3 1 2>&1 | grep -E "it 0:|total|crc" | sed -e 's/refit.*seeds/seeds/' -e 's/shift.*//' | tr '\n' ' '; echo; done
