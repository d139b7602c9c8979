set -x
cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-prove"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["phases_ms_per_step"].items()}, d["root"][:2])'
timeout 600 python -m pytest tests/test_gpu_commit_parity.py -x -q -m gpu > gpurun_out/r2_c16_parity_tma.log 2>&1; echo "parity tma rc=$?"
tail -n 5 gpurun_out/r2_c16_parity_tma.log
for v in "BFGPU_NTT_TMA=1 BFGPU_NTT_TURN=1" "BFGPU_NTT_TMA=0 BFGPU_NTT_TURN=0"; do
  echo "== $v"
  env $v timeout 300 $B 2>gpurun_out/r2_c16_err.txt | python -c "$S"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_pass|k_ingest" -c 12 --csv --log-file gpurun_out/r2_c16_lde_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-prove > gpurun_out/r2_c16_ncu.log 2>&1; echo "ncu rc=$?"
