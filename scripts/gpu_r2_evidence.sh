#!/bin/bash
# round 2 evidence: parity tests, contract bench line (both arms), ncu launch lists, ncu --set full of the dominant kernel and of
# the silence-anchoring chain, phase timers, compute-sanitizer.  Everything lands in gpurun_out/; scripts/save_profiles_r2.py
# turns it into profiles/r02_<tag>_*.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
# launch lists: the contract command, then the other workloads
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ncu_launch_run.log 2>&1
for w in sil 3 4; do
  STEPS=2 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv python scripts/variant_launches.py $w > gpurun_out/variant_$w.log 2>&1
done
python scripts/variant_launches.py > gpurun_out/variants_plain.log 2>&1; cat gpurun_out/variants_plain.log
# ncu --set full: the dominant kernel on the metric batch, the kernels of the silence-anchoring chain
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:band3_direct_kernel" -s 6 -c 1 -f -o gpurun_out/prof_direct \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ncu_full_run.log 2>&1
bash scripts/gpu_ncu_sil.sh sil
ls -la gpurun_out/*.ncu-rep
if [ -f bournemouth-forced-aligner_b200/lib/libbfa_b200_prof.so ]; then
BFA_B200_LIB=$PWD/bournemouth-forced-aligner_b200/lib/libbfa_b200_prof.so timeout 300 python scripts/phase_prof.py 4096 > gpurun_out/phase.log 2>&1
head -20 gpurun_out/phase.log
fi
bash scripts/gpu_sanitize.sh 2>&1 | tail -12
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
