cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_dist_commit.py tests/test_gpu_dist_prove.py -x -q -m gpu > gpurun_out/r2_c38_parity.log 2>&1; echo "parity rc=$?"
tail -n 6 gpurun_out/r2_c38_parity.log
for fs in 1 0; do
BFGPU_DIST_FUSED_SCATTER=$fs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$fs bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-prove --no-e2e 2> gpurun_out/r2_c38_err_$fs.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fused_scatter=$fs', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['root_matches_oracle_golden'])"
done
