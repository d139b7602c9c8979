set -x
cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_commit_parity.py tests/test_gpu_golden_vectors.py -x -q -m gpu > gpurun_out/r2_c25_parity.log 2>&1; echo "parity rc=$?"
tail -n 12 gpurun_out/r2_c25_parity.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-prove --no-e2e"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["phases_ms_per_step"].items()}, d["root"][:2], d.get("e2e"))'
for v in "BFGPU_NTT_CFWD=1" "BFGPU_NTT_CFWD=0"; do
  echo "== $v"
  env $v timeout 300 $B 2>gpurun_out/r2_c25_err.txt | python -c "$S"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_pass|k_ingest|k_cfwd" -c 10 --csv --log-file gpurun_out/r2_c25_lde_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-prove > gpurun_out/r2_c25_ncu.log 2>&1; echo "ncu rc=$?"
