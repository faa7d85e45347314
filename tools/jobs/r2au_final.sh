python -m pytest tests -m gpu -x -q > gpurun_out/r2au_pytest.log 2>&1; tail -3 gpurun_out/r2au_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2au_bench.json 2> gpurun_out/r2au_bench.err; echo bench rc=$?; tail -c 300 gpurun_out/r2au_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2au_bench.json'))
print('value %.4e asm %.3f kern %.3f step %.2f e2e %.3f frac %.3f'%(d['value'], d['assembly_stage_ms'],d['assembly_kernel_ms'],d['ms_per_step'],d['e2e']['ms_per_step'], d['roofline']['frac']), d['parity']['ok'], d['clocks'])
for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if not isinstance(b,(dict,str))})
print(d.get('cpu_baseline'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2au_launches.csv python bench.py --steps 2 --warmup 1 --no-extra-configs --no-cpu-baseline --no-parity > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_fluid_tet4_grouped --launch-skip 1 -c 1 -f -o gpurun_out/r2au_asm python tools/prof_assemble.py 118 120 2 > /dev/null 2>&1
