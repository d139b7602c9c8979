set -x
cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c43_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/r2_c43_pytest.log
timeout 900 python bench.py > gpurun_out/r2_c43_bench.json 2> gpurun_out/r2_c43_bench.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c43_smoke.log 2>&1; echo "smoke rc=$?"
