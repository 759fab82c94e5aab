#!/bin/bash
# tools/build_variant.sh NAME "EXTRA NVCC FLAGS": libnanowrap_NAME.so with sweep.cu (and tree.cu) compiled with extra flags;
# use it with NANOWRAP_LIB=ch_shrinkwrap_b200/libnanowrap_NAME.so (kernel-variant A/B measurements)
set -e
cd "$(dirname "$0")/../ch_shrinkwrap_b200"
python -m ch_shrinkwrap_b200.build >/dev/null 2>&1 || (cd .. && python -m ch_shrinkwrap_b200.build >/dev/null)
mkdir -p build/var_$1
for f in sweep tree api mesh_ops; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c csrc/$f.cu -o build/var_$1/$f.cu.o &
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -fmad=false $2 -c csrc/curvature.cu -o build/var_$1/curvature.cu.o &
wait
objs=""
for o in build/*.cu.o; do b=$(basename $o); if [ -f build/var_$1/$b ]; then objs="$objs build/var_$1/$b"; else objs="$objs $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libnanowrap_$1.so $objs -ldl
echo ch_shrinkwrap_b200/libnanowrap_$1.so
