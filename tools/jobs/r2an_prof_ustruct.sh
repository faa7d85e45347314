ncu --set full --clock-control none --import-source on -k regex:assemble_ustruct_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r2an_us8 python tools/prof_ustruct_hex8.py 80 > /dev/null 2>&1
python tools/ncu_hot.py gpurun_out/r2an_us8.ncu-rep 40 > gpurun_out/r2an_us8_hot.txt 2>&1; head -75 gpurun_out/r2an_us8_hot.txt
