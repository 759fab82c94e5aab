cd $GRAFT_REPO_ROOT
echo "== default: 256(a,b) + 128(d)"; python tools/trace_probe.py c3 2 2>&1 | grep -E "it 3|total|crc"
echo "== ldg128"; NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap_ldg128.so python tools/trace_probe.py c3 2 2>&1 | grep -E "it 3|total|crc"
echo "== default again"; python tools/trace_probe.py c3 2 2>&1 | grep -E "it 3|total|crc"
