cd $GRAFT_REPO_ROOT
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --workload c4s --no-cpu-baseline ) > gpurun_out/r2_bench_c4s_n8_v3.json 2> gpurun_out/r2_bench_c4s_n8_v3.err; tail -c 900 gpurun_out/r2_bench_c4s_n8_v3.json; tail -4 gpurun_out/r2_bench_c4s_n8_v3.err
