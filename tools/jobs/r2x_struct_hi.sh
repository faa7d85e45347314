python -m pytest tests/test_gpu_struct.py -x -q -m gpu 2>&1 | tail -8
python tools/bench_struct.py 171 3 2>&1 | tail -4
