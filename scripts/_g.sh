mkdir -p gpurun_out
(timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "reverb or conv or timed or plugin or swap or biquad or chunk or graph or port" 2>&1 | tail -12) > gpurun_out/r02_memcheck.log 2>&1; tail -6 gpurun_out/r02_memcheck.log | cut -c1-200
(timeout 700 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "parity or timed or swap or generic" 2>&1 | tail -12) > gpurun_out/r02_racecheck.log 2>&1; tail -6 gpurun_out/r02_racecheck.log | cut -c1-200
