cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_commit_parity.py tests/test_gpu_golden_vectors.py tests/test_gpu_determinism.py -x -q -m gpu > gpurun_out/r2_c46_parity.log 2>&1; echo "parity rc=$?"
tail -n 3 gpurun_out/r2_c46_parity.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-prove --no-e2e"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["phases_ms_per_step"].items()}, d["root_matches_oracle_golden"])'
timeout 300 $B 2>>gpurun_out/r2_c46_err.txt | python -c "$S"
for v in "BFGPU_NTT_TMA=0" "BFGPU_NTT_TMA=0 BFGPU_NTT_TURN=0"; do echo "== $v"; env $v timeout 300 $B 2>>gpurun_out/r2_c46_err.txt | python -c "$S"; done
