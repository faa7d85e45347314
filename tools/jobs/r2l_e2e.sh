python -m pytest tests/test_gpu_fluid.py -x -q -m gpu 2>&1 | tail -3
python bench.py --no-extra-configs --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo rc=$?; tail -c 600 gpurun_out/r2l_bench.err
SVB200_HOST_NO_PIPELINE=1 python bench.py --no-extra-configs --no-cpu-baseline --no-parity > gpurun_out/r2l_bench_nopipe.json 2>/dev/null; echo rc=$?
python - <<'PY'
import json
for f in ('r2l_bench','r2l_bench_nopipe'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, 'asm %.3f kern %.3f e2e %.3f ms, e2e value %.4e'%(d['assembly_stage_ms'],d['assembly_kernel_ms'],d['e2e']['ms_per_step'],d['e2e']['value']), d['e2e'].get('R_max_rel_vs_plain_sequence'), d.get('parity') and d['parity']['ok'])
PY
