LD_PRELOAD=$PWD/tools/debug/segv_trace.so python -m pytest -p no:faulthandler "tests/test_gpu_hostshim.py::test_struct_newton_iteration_through_cpp_plugin" -x -q -m gpu > gpurun_out/r2s_pytest.log 2>&1
grep -A40 "native backtrace" gpurun_out/r2s_pytest.log | head -60
