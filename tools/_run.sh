cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_reference_c3.json 2> gpurun_out/r2_bench_reference_c3.err; tail -c 1200 gpurun_out/r2_bench_reference_c3.json; tail -5 gpurun_out/r2_bench_reference_c3.err
