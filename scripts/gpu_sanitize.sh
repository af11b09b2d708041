#!/bin/bash
# compute-sanitizer over a slice of the GPU parity tests: memcheck, racecheck (shared memory), synccheck
mkdir -p gpurun_out
SEL="viterbi_core_golden or decode_forced_golden or packed_rows_on_every_alignment or segmented_packed or long_targets_wide or test_process_sentence_api_vs_reference or stitch_log_softmax or alignment_score or logits_in_one_kernel or downstream_steps_from_logits or takes_the_logits_path or logits_through_the_planner_chain or decode_alignments_from_logits"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log \
   python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/memcheck_pytest.log; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -3 gpurun_out/memcheck.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/synccheck.log \
   python -m pytest tests -m gpu -x -q -k "decode_forced_golden or segmented_packed or long_targets_wide or test_process_sentence_api_vs_reference or logits_in_one_kernel or takes_the_logits_path or logits_through_the_planner_chain" > gpurun_out/synccheck_pytest.log 2>&1
echo "synccheck rc=$?"; tail -2 gpurun_out/synccheck_pytest.log; tail -3 gpurun_out/synccheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --log-file gpurun_out/racecheck.log \
   python -m pytest tests -m gpu -x -q -k "decode_forced_golden or segmented_packed or long_targets_wide" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/racecheck_pytest.log
grep "Race reported" gpurun_out/racecheck.log | sed -e 's/.*Race reported between //' -e 's/ and .*//' | sort | uniq -c | sort -rn | head -30
tail -3 gpurun_out/racecheck.log
