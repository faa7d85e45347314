python -m pytest tests/test_gpu_fluid.py tests/test_gpu_hostshim.py -x -q -m gpu -k "general_element or quadratic or hex8" 2>&1 | tail -3
python tools/bench_quadratic.py 36 28 2>&1 | grep fluid
python tools/bench_fluid_hex8.py 2>&1 | tail -3
