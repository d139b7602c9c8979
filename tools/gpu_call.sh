cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_prove_parity.py tests/test_gpu_open_parity.py tests/test_gpu_dist_prove.py tests/test_gpu_tracegen_parity.py tests/test_gpu_plug_point2.py -x -q -m gpu > gpurun_out/r2_c31_parity.log 2>&1; echo "parity rc=$?"
tail -n 4 gpurun_out/r2_c31_parity.log
for pgm in loop22 loop20 fibo hello; do
timeout 300 python tools/prove_phases.py $pgm 4 2>>gpurun_out/r2_c31_err.txt | tee -a gpurun_out/r2_c31_phases.jsonl
done
python - <<'PY'
import time, sys
sys.path.insert(0,'.')
import zkvm_brainfuck_b200 as bf
ctx=bf.Context(0)
code="++++++++[>-[>-[>+>+<<-]<-]<-]"
for _ in range(5):
    t=time.perf_counter(); rec=bf.Record(code, ctx=ctx); dt=time.perf_counter()-t
    print('executor', round(dt*1e3,2), 'ms', rec.cycles, round(dt*1e9/rec.cycles,2),'ns/cycle'); rec.free()
PY
