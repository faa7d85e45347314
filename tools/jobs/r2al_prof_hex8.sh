ncu --set full --clock-control none --import-source on -k regex:assemble_fluid_gen --launch-skip 1 -c 1 -f -o gpurun_out/r2al_fg8 python tools/prof_fluid_hex8.py 100 > /dev/null 2>&1
python tools/ncu_hot.py gpurun_out/r2al_fg8.ncu-rep 30 > gpurun_out/r2al_fg8_hot.txt 2>&1; head -60 gpurun_out/r2al_fg8_hot.txt
python tools/bench_phys.py 80 100 3 2>&1 | tee gpurun_out/r2al_phys.txt
