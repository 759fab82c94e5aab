cd $GRAFT_REPO_ROOT
echo "== scalar, new layout"; python tools/trace_probe.py c3 2 2>&1 | grep -E "it |total|crc"
echo "== packed"; NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap_packed.so python tools/trace_probe.py c3 2 2>&1 | grep -E "it |total|crc"
