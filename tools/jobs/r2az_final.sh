python -m pytest tests -m gpu -x -q > gpurun_out/r2az_pytest.log 2>&1; tail -2 gpurun_out/r2az_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py s3 2>&1 | grep -E "RACECHECK SUMMARY|sanitize_small" > gpurun_out/r2az_racecheck.txt; cat gpurun_out/r2az_racecheck.txt
