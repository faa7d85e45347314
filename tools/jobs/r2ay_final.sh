python -m pytest tests -m gpu -x -q > gpurun_out/r2ay_pytest.log 2>&1; tail -3 gpurun_out/r2ay_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2ay_bench.json 2> gpurun_out/r2ay_bench.err; echo bench rc=$?; tail -c 300 gpurun_out/r2ay_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ay_bench.json'))
print('value %.4e asm %.3f kern %.3f step %.2f e2e %.3f frac %.3f'%(d['value'], d['assembly_stage_ms'],d['assembly_kernel_ms'],d['ms_per_step'],d['e2e']['ms_per_step'], d['roofline']['frac']), d['parity']['ok'], d['clocks'])
for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if not isinstance(b,(dict,str))})
print(d.get('cpu_baseline'))
PY
