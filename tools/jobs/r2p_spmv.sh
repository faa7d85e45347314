python -m pytest tests/test_gpu_spmv_variants.py -x -q -m gpu 2>&1 | tail -5
python tools/spmv_variants.py both 5 > gpurun_out/r2p_spmv_variants.txt 2>&1; cat gpurun_out/r2p_spmv_variants.txt
python -m pytest tests/test_gpu_fluid.py tests/test_gpu_struct.py tests/test_gpu_heat.py tests/test_gpu_solver_equality.py -x -q -m gpu 2>&1 | tail -3
python tools/bench_ns.py 2>&1 | tail -2
SVB200_SCHUR_UNFUSED=1 python tools/bench_ns.py 2>&1 | tail -1
