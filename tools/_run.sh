cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_12.txt 2>&1; tail -5 gpurun_out/r2_gputest_12.txt
echo "== trace"; python tools/trace_probe.py c3 2 2>&1 | grep -E "it |total|crc"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_adj2.json 2> gpurun_out/r2_bench_c3_adj2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3_adj2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['stages']); print({k:(v.get('ms'),v.get('frac')) for k,v in d['kernels'].items()})"
