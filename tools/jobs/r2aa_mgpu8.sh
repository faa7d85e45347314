TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563"
$TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2aa_n8.json 2> gpurun_out/r2aa_n8.err; echo rc=$?; tail -c 800 gpurun_out/r2aa_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2aa_n8.json'))
print('value %.4e asm %.3f kern %.3f step %.1f it %.4f e2e %.3f'%(d['value'], d['assembly_stage_ms'], d['assembly_kernel_ms'], d['ms_per_step'], d['gmres']['ms_per_iteration'], d['e2e']['ms_per_step']))
print(d['e2e']['copies_alone_ms']); print(d['e2e']['device_timeline_ms_rank0']); print(d['run']['numa'])
print('parity', d['parity']['ok'], d['parity']['assembly_max_rel'], d['parity'].get('c5_fsi_R_max_rel'))
print(json.dumps(d.get('configs'))[:1200])
PY
(nvidia-smi topo -m | head -12; lscpu | grep -i -E "numa|socket|^cpu\(s\)") > gpurun_out/r2aa_topo.txt 2>&1
