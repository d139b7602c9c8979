cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-prove"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["value"],1), d["root_matches_oracle_golden"])'
for sch in "" "64,64,64,40,24" "72,72,64,48" "80,80,64,32" "64,64,48,48,32" "96,64,48,32,16" "64,64,64,32,16,16"; do
  echo "== schedule '$sch'"; BFGPU_PIPE_SCHEDULE=$sch timeout 300 $B 2>>gpurun_out/r2_c42_err.txt | python -c "$S"
done
