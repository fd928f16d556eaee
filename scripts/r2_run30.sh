# round 2, GPU call 30: compute-sanitizer (memcheck, then racecheck / initcheck on a smaller case) over the pipeline parity tests
set -x
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "test_pipeline_parity or ragged or stages_compose or index_matches" > gpurun_out/r2ac_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/r2ac_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "test_pipeline_parity and 100-0" > gpurun_out/r2ac_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -12 gpurun_out/r2ac_racecheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "test_pipeline_parity and 100-0" > gpurun_out/r2ac_initcheck.log 2>&1
echo "initcheck rc=$?"; tail -12 gpurun_out/r2ac_initcheck.log
