TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558"
$TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n_n8_blocks.json 2> gpurun_out/r2n_n8_blocks.err; echo blocks rc=$?
tail -c 600 gpurun_out/r2n_n8_blocks.err
SVB200_HALO_PER_NEIGHBOUR=1 SVB200_DOT_NO_FUSE=1 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-parity > gpurun_out/r2n_n8_blocks_old.json 2> gpurun_out/r2n_n8_blocks_old.err; echo old rc=$?
$TR bench.py --gpus 8 --steps 5 --warmup 3 --partition slab --no-parity > gpurun_out/r2n_n8_slab.json 2> gpurun_out/r2n_n8_slab.err; echo slab rc=$?
