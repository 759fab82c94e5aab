cd $GRAFT_REPO_ROOT
python tools/kernel_probe.py curvature
for v in cb64 cb32 cb256; do echo -n "$v: "; NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap_$v.so python tools/kernel_probe.py curvature; done
