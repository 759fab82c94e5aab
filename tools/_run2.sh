cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -v 2>&1 | tail -12 > gpurun_out/r2_gputest_multi_2gpu_v2.txt; cat gpurun_out/r2_gputest_multi_2gpu_v2.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2_bench_c3_n2_v2.json 2> gpurun_out/r2_bench_c3_n2_v2.err; tail -c 700 gpurun_out/r2_bench_c3_n2_v2.json
