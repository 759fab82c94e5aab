cd $GRAFT_REPO_ROOT
for w in c3 c5; do python tools/kprobe_wl.py $w mesh_prior; for v in 2 4 5; do echo -n "minb$v: "; NANOWRAP_LIB=$PWD/ch_shrinkwrap_b200/libnanowrap_mp$v.so python tools/kprobe_wl.py $w mesh_prior; done; done
