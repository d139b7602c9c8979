cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-prove --no-e2e"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["phases_ms_per_step"].items()}, d["root_matches_oracle_golden"])'
echo "== default"; timeout 300 $B 2>>gpurun_out/r2_c41_err.txt | python -c "$S"
for so in variants/libbfgpu_p2_e2_i1.so variants/libbfgpu_p2_e4_i1.so variants/libbfgpu_p2_e1_i13.so variants/libbfgpu_p2_e2_i13.so; do
  echo "== $so"; BFGPU_SO=$GRAFT_REPO_ROOT/$so timeout 300 $B 2>>gpurun_out/r2_c41_err.txt | python -c "$S"
done
