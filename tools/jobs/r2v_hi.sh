python -m pytest tests/test_gpu_fluid.py -x -q -m gpu -k "quadratic or general_element" 2>&1 | tail -8
python -m pytest tests/test_gpu_hostshim.py -x -q -m gpu 2>&1 | tail -8
