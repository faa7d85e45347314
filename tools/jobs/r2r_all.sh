python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
python tools/bench_ns.py 2>&1 | tail -1
SVB200_SCHUR_UNFUSED=1 python tools/bench_ns.py 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo rc=$?; tail -c 400 gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print('asm %.3f kern %.3f step %.2f e2e %.3f'%(d['assembly_stage_ms'],d['assembly_kernel_ms'],d['ms_per_step'],d['e2e']['ms_per_step']), d['parity']['ok'])
for k,v in d.get('configs',{}).items(): print(k, json.dumps(v)[:900])
PY
