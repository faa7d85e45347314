TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --partition blocks > gpurun_out/r2ax_n2.json 2> gpurun_out/r2ax_n2.err; echo rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/r2ax_n2.json'))
print('value %.4e asm %.3f kern %.3f step %.2f e2e %.3f parity %s itr %d'%(d['value'], d['assembly_stage_ms'], d['assembly_kernel_ms'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity']['ok'], d['gmres']['itr']), d['run']['partition'][:40], d['gmres'].get('ms_per_iteration'))
for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if not isinstance(b,(dict,str))})
PY
