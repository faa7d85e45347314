python -m pytest tests/test_gpu_struct.py tests/test_gpu_hostshim.py -x -q -m gpu 2>&1 | tail -15
