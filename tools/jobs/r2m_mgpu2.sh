TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557"
$TR bench.py --gpus 2 --steps 3 --warmup 3 --partition blocks > gpurun_out/r2m_n2_blocks.json 2> gpurun_out/r2m_n2_blocks.err; echo blocks rc=$?
tail -c 600 gpurun_out/r2m_n2_blocks.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_n2_blocks.json'))
print('asm %.3f e2e %.3f ms'%(d['assembly_stage_ms'],d['e2e']['ms_per_step']), d['e2e'].get('R_max_rel_vs_plain_sequence'), d['parity']['ok'], d['gmres'])
PY
python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "slab and gmres" 2>&1 | tail -2
