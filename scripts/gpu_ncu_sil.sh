#!/bin/bash
# ncu --set full of the planner chain kernels on the SIL-bearing metric shape (one launch each)
mkdir -p gpurun_out
W=${1:-sil}
STEPS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:silprob_kernel|plan_kernel|viterbi_band3_kernel|assort_confidence_kernel|silunits_kernel" -c 5 -f -o gpurun_out/prof_$W \
  python scripts/variant_launches.py $W > gpurun_out/ncu_$W.log 2>&1
tail -3 gpurun_out/ncu_$W.log
ls -la gpurun_out/prof_$W.ncu-rep
