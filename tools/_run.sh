cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_14.txt 2>&1; tail -3 gpurun_out/r2_gputest_14.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_v4.json 2> gpurun_out/r2_bench_c3_v4.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3_v4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
ncu --nvtx --nvtx-include "adjoint/" --metrics gpu__time_duration.sum --clock-control none -c 4 python tools/nvtx_probe.py 2>&1 | grep -E "k_adjoint|gpu__time|No kernels" | head -6
