cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c40_bench_n4.json 2> gpurun_out/r2_c40_bench_n4.err; echo "bench n4 rc=$?"
