#!/bin/bash
# compute-sanitizer over a small slice of the GPU parity tests (memcheck, then racecheck on shared memory)
mkdir -p gpurun_out
export BFA_SANITIZE=1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log \
   python -m pytest tests -m gpu -x -q -k "viterbi_core_golden or decode_forced_golden or packed_rows_on_every_alignment or segmented_packed" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/memcheck_pytest.log; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --log-file gpurun_out/racecheck.log \
   python -m pytest tests -m gpu -x -q -k "decode_forced_golden or segmented_packed" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/racecheck_pytest.log; tail -12 gpurun_out/racecheck.log
