#!/bin/bash
# interleaved A/B: REPS rounds of scripts/quick_time.py over the default library ("default") and lib/libbfa_b200_<variant>.so;
# a variant prefixed with "t:" first runs a subset of the parity tests against that build
mkdir -p gpurun_out
REPS=${REPS:-3}
vs=()
for v in "$@"; do
  case $v in t:*) v=${v#t:}
    BFA_B200_LIB=$PWD/bournemouth-forced-aligner_b200/lib/libbfa_b200_$v.so timeout 600 python -m pytest tests/test_cuda_full.py tests/test_cuda_parity.py -m gpu -x -q -k "metric or ragged or random or pipelined or config2 or config4 or flip" 2>&1 | tail -3;;
  esac
  vs+=($v)
done
for r in $(seq 1 $REPS); do
  for v in "${vs[@]}"; do
    if [ $v == default ]; then python scripts/quick_time.py; else BFA_B200_LIB=$PWD/bournemouth-forced-aligner_b200/lib/libbfa_b200_$v.so python scripts/quick_time.py; fi
  done
done
