TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
for v in zc nozc; do
  if [ $v = nozc ]; then export SVB200_HOST_NO_ZEROCOPY=1; fi
  $TR bench.py --gpus 2 --steps 5 --warmup 3 --partition blocks --no-parity > gpurun_out/r2w_n2_$v.json 2> gpurun_out/r2w_n2.err; echo $v rc=$?
  python - <<PY
import json
d=json.load(open('gpurun_out/r2w_n2_$v.json'))
print('$v asm %.3f kern %.3f e2e %.3f err %.1e'%(d['assembly_stage_ms'], d['assembly_kernel_ms'], d['e2e']['ms_per_step'], d['e2e']['R_max_rel_vs_plain_sequence']), d['e2e']['device_timeline_ms_rank0'])
PY
done
unset SVB200_HOST_NO_ZEROCOPY
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
