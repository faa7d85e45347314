python tools/bench_struct_tet4.py 2>&1 | tail -3
python tools/bench_phys.py 50 100 5 ustruct 2>&1 | tail -3
python tools/bench_fsi.py 2>&1 | tail -8
python tools/bench_struct.py 171 3 2>&1 | grep assemble | tail -1
python -m pytest tests/test_gpu_struct.py tests/test_gpu_ustruct.py -x -q -m gpu 2>&1 | tail -2
