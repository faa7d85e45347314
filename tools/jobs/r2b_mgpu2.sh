python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2b_pytest_mgpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2b_pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
$TR bench.py --gpus 2 --steps 3 --warmup 3 --partition blocks > gpurun_out/r2b_n2_blocks.json 2> gpurun_out/r2b_n2_blocks.err; echo blocks rc=$?
SVB200_HALO_PER_NEIGHBOUR=1 SVB200_DOT_NO_FUSE=1 $TR bench.py --gpus 2 --steps 3 --warmup 3 --partition blocks --no-parity > gpurun_out/r2b_n2_blocks_old.json 2> gpurun_out/r2b_n2_blocks_old.err; echo old rc=$?
$TR bench.py --gpus 2 --steps 3 --warmup 3 --partition slab > gpurun_out/r2b_n2_slab.json 2> gpurun_out/r2b_n2_slab.err; echo slab rc=$?
tail -c 1500 gpurun_out/r2b_n2_blocks.err
