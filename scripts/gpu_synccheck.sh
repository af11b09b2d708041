#!/bin/bash
mkdir -p gpurun_out
for K in "test_metric_shape_vs_oracle" "decode_forced_golden" "segmented_packed" "test_process_sentence_api_vs_reference"; do
timeout 900 compute-sanitizer --tool synccheck --print-limit 3 --error-exitcode 7 --log-file gpurun_out/synccheck_$K.log \
   python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/synccheck_pytest_$K.log 2>&1
echo "== $K synccheck rc=$?"; tail -1 gpurun_out/synccheck_pytest_$K.log; grep "Barrier error\|    at bfa\|by thread\|Device Frame" gpurun_out/synccheck_$K.log | head -8; tail -1 gpurun_out/synccheck_$K.log
done
