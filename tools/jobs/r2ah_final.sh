python -m pytest tests -m gpu -x -q > gpurun_out/r2ah_pytest.log 2>&1; tail -3 gpurun_out/r2ah_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err; echo bench rc=$?; tail -c 200 gpurun_out/r2ah_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ah_bench.json'))
print('value %.4e asm %.3f kern %.3f step %.2f e2e %.3f frac %.3f'%(d['value'], d['assembly_stage_ms'],d['assembly_kernel_ms'],d['ms_per_step'],d['e2e']['ms_per_step'], d['roofline']['frac']), d['parity']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
for k,v in d.get('configs',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,(dict,str))})
PY
