run() { env "$@" python bench.py --no-extra-configs --no-cpu-baseline --no-parity --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', 'asm %.3f kern %.3f value %.4e'%(d['assembly_stage_ms'],d['assembly_kernel_ms'],d['value']), d['clocks']['sm_mhz'])"; }
run SVB200_EAGER_ZERO=1
run X=1
run SVB200_ZERO_THREADS=64
run SVB200_ZERO_THREADS=128
run SVB200_ZERO_THREADS=128 SVB200_ZERO_CTAS_PER_SM=2
run SVB200_ZERO_CHUNK0_DIV=4
run SVB200_ZERO_CHUNK0_DIV=3 SVB200_ZERO_THREADS=64
run SVB200_EAGER_ZERO=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 1 --warmup 1 --no-extra-configs --no-cpu-baseline --no-parity > /dev/null 2>&1
grep -E "zero_small|assemble_fluid_tet4_grouped" gpurun_out/r2k_launches.csv | head -12
