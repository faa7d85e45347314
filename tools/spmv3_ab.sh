python -m pytest tests/test_gpu_struct.py tests/test_gpu_fullsize_struct.py -q -m gpu 2>&1 | tail -3
for mode in thread warp; do SVB200_SPMV3=$mode python tools/bench_struct.py 171 1 2>&1 | grep "SpMV\|BiCG" | sed "s/^/$mode: /"; done
