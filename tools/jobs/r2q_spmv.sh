python -m pytest tests/test_gpu_spmv_variants.py -x -q -m gpu 2>&1 | tail -3
python tools/spmv_variants.py both 5 > gpurun_out/r2q_spmv_variants.txt 2>&1; cat gpurun_out/r2q_spmv_variants.txt
for i in 1 2 3; do python -m pytest "tests/test_gpu_fluid.py::test_capped_coupled_face_parity" -x -q -m gpu -s 2>&1 | grep -E "cap:|assert|Error|passed|failed" | head -8; done
ncu --set full --clock-control none -k regex:"spmv|schur" -f -o gpurun_out/r2q_spmv python tools/spmv_variants.py prof > gpurun_out/r2q_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/r2q_spmv.ncu-rep > gpurun_out/r2q_spmv_table.txt 2>&1; cat gpurun_out/r2q_spmv_table.txt
