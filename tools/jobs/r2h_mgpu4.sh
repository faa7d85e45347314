# 4 GPUs: multi-GPU parity tests (scattered / FSI / METIS partitions on 4 ranks, nodes shared by more than two ranks), then
# bench.py on the 2x2x1 block partition with the fused exchange, the round-1 per-neighbour exchange, and on slabs
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2h_pytest_mgpu4.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2h_pytest_mgpu4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556"
$TR bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2h_n4_blocks.json 2> gpurun_out/r2h_n4_blocks.err; echo blocks rc=$?
SVB200_HALO_PER_NEIGHBOUR=1 SVB200_DOT_NO_FUSE=1 $TR bench.py --gpus 4 --steps 3 --warmup 3 --no-parity > gpurun_out/r2h_n4_blocks_old.json 2> gpurun_out/r2h_n4_blocks_old.err; echo old rc=$?
SVB200_HALO_NO_OVERLAP=1 $TR bench.py --gpus 4 --steps 3 --warmup 3 --no-parity > gpurun_out/r2h_n4_blocks_noovl.json 2> gpurun_out/r2h_n4_blocks_noovl.err; echo noovl rc=$?
$TR bench.py --gpus 4 --steps 3 --warmup 3 --partition slab --no-parity > gpurun_out/r2h_n4_slab.json 2> gpurun_out/r2h_n4_slab.err; echo slab rc=$?
tail -c 800 gpurun_out/r2h_n4_blocks.err
