ncu --set full --clock-control none --import-source on -k regex:assemble_ustruct_tet4 --launch-skip 1 -c 1 -f -o gpurun_out/r2ag_ustruct_tet4 python tools/bench_phys.py 50 100 1 ustruct > /dev/null 2>&1
python tools/ncu_hot.py gpurun_out/r2ag_ustruct_tet4.ncu-rep 30 > gpurun_out/r2ag_ustruct_tet4_hot.txt 2>&1; head -70 gpurun_out/r2ag_ustruct_tet4_hot.txt
