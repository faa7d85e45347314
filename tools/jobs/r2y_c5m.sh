TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562"
$TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2y_n2.json 2> gpurun_out/r2y_n2.err; echo rc=$?; tail -c 1500 gpurun_out/r2y_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2y_n2.json'))
print('asm %.3f e2e %.3f parity'%(d['assembly_stage_ms'], d['e2e']['ms_per_step']), d['parity']['ok'], d['parity'].get('c5_fsi_R_max_rel'))
print(json.dumps(d.get('configs'))[:1500])
PY
