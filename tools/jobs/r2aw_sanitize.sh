echo "# compute-sanitizer over the round-2 session-3 kernels (tools/sanitize_small.py s3): FSI-ustruct + masked ustruct_r, URIS (split launch + general kernel), RIS take/put/apply, heat / lElas / mesh on quadratic + wedge elements, HEX8 heat / lElas lane-per-Gauss-point kernels, rolled ustruct TET4, Taylor-Hood fluid + thood_val_rc, degenerate sizes" > gpurun_out/r2aw_compute_sanitizer.txt
echo "## memcheck" >> gpurun_out/r2aw_compute_sanitizer.txt
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_small.py s3 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|sanitize_small|Invalid|Error|error" | head -30 >> gpurun_out/r2aw_compute_sanitizer.txt
echo "## racecheck" >> gpurun_out/r2aw_compute_sanitizer.txt
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_small.py s3 2>&1 | grep -E "COMPUTE-SANITIZER|RACECHECK SUMMARY|sanitize_small|hazard|Error|error" | head -30 >> gpurun_out/r2aw_compute_sanitizer.txt
cat gpurun_out/r2aw_compute_sanitizer.txt
