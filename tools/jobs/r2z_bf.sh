python -m pytest tests/test_gpu_fluid.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for v in vector always vector; do
  unset SVB200_BF_ALWAYS
  if [ $v = always ]; then export SVB200_BF_ALWAYS=1; fi
  python tools/ab_assemble.py 118 120 10 2>&1 | tail -1 | sed "s/^/$v /"
done
