cd $GRAFT_REPO_ROOT
for r in 0 8 12 16 24 32 64; do echo -n "NW_SEED_WIDE=$r: "; NW_SEED_WIDE=$r python tools/trace_probe.py c3 1 2>&1 | grep -E "it 0:|total|crc" | sed -e 's/refit.*seeds/seeds/' -e 's/shift.*//' | tr '\n' ' '; echo; done
