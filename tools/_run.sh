cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_13.txt 2>&1; tail -3 gpurun_out/r2_gputest_13.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_v3.json 2> gpurun_out/r2_bench_c3_v3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3_v3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']); print({k:round(v['ms_total']/d['steps'],3) for k,v in d['stages'].items()}); print({k:(round(v.get('ms',0),4),round(v.get('frac',0),3)) for k,v in d['kernels'].items()})"
