cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_15.txt 2>&1; tail -3 gpurun_out/r2_gputest_15.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_v5.json 2> gpurun_out/r2_bench_c3_v5.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3_v5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']); print({k:round(v['ms_total']/d['steps'],3) for k,v in d['stages'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_c3_v3.csv python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
