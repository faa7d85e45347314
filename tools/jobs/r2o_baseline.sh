# round-2 session-2 baseline: GPU tests at HEAD, full ncu capture of the TET4 fluid kernel, NS launch list, host topology
python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
(nvidia-smi topo -m; lscpu | grep -i -E "numa|model name|socket|^cpu\(s\)"; echo cpuset $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null); nproc; free -g | head -2) > gpurun_out/r2o_topo.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_fluid_tet4_grouped --launch-skip 1 -c 1 -f -o gpurun_out/r2o_asm python tools/prof_assemble.py 118 120 2 > gpurun_out/r2o_ncu_asm.log 2>&1
python tools/ncu_hot.py gpurun_out/r2o_asm.ncu-rep 40 > gpurun_out/r2o_asm_hot.txt 2>&1; head -30 gpurun_out/r2o_asm_hot.txt
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 6000 --csv --log-file gpurun_out/r2o_ns_launches.csv python tools/bench_ns.py > gpurun_out/r2o_ns.log 2>&1
tail -3 gpurun_out/r2o_ns.log
ls -la gpurun_out/ | tail -8
