mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/r2v_plain.log 2>&1; tail -3 gpurun_out/r2v_plain.log
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 3000 > gpurun_out/r2v_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "^ok" gpurun_out/r2v_memcheck.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r2v_memcheck.log | head -10
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py 2000 > gpurun_out/r2v_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "^ok" gpurun_out/r2v_racecheck.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r2v_racecheck.log | head -10
