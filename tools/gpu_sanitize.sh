mkdir -p gpurun_out
( timeout 10 tools/sanitize/driver 1500 4000 6000; echo "plain rc=$?" ) > gpurun_out/sanitize.log 2>&1
( timeout 25 compute-sanitizer --tool memcheck --print-limit 5 tools/sanitize/driver 1500 3000 3000 2>&1 | tail -8; echo "memcheck done" ) >> gpurun_out/sanitize.log 2>&1
( timeout 25 compute-sanitizer --tool racecheck --print-limit 5 tools/sanitize/driver 1500 3000 3000 2>&1 | tail -8; echo "racecheck done" ) >> gpurun_out/sanitize.log 2>&1
cat gpurun_out/sanitize.log
