cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1200 python tools/stress_lde.py 12 > gpurun_out/r2_c36_stress.log 2>&1; echo "stress rc=$?"
tail -n 12 gpurun_out/r2_c36_stress.log
