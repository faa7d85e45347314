ncu --set full --clock-control none --import-source on -k regex:assemble_fluid_gen --launch-skip 1 -c 1 -f -o gpurun_out/r2aq_t10 python tools/prof_fluid_tet10.py 36 > /dev/null 2>&1
python tools/ncu_hot.py gpurun_out/r2aq_t10.ncu-rep 30 > gpurun_out/r2aq_t10_hot.txt 2>&1; head -64 gpurun_out/r2aq_t10_hot.txt
