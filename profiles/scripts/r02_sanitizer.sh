# compute-sanitizer memcheck of the row kernels and one config-1 train step (B200 box, repo root): bash profiles/scripts/r02_sanitizer.sh
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 30 python tests/gpu_sanitizer_step.py > gpurun_out/r02_sanitizer.log 2>&1
echo "sanitizer rc=$?"
grep -E "ERROR SUMMARY|done \(|Invalid|out of bounds|Error" gpurun_out/r02_sanitizer.log | head -20
tail -3 gpurun_out/r02_sanitizer.log
