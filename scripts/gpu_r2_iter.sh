#!/bin/bash
# round 2 iteration: parity tests + device-timed bench variants + phase timers (if the prof build exists) (+ optional extra command)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
show() { python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:600]); continue
    r = d['roofline']
    print('$1: value %.3e ms/step %.4f kernel_ms %.4f frac %.3f isolated %.4f fill %.4f launches %d' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r.get('isolated_launch_ms', 0), r.get('fill_phase', {}).get('kernel_ms', 0), d['gpu_launches']))
"; }
for v in "" "--no-pipeline" "--chain"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline $v > gpurun_out/bench_quick$v.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_quick$v.log
  tail -3 gpurun_out/bench_quick$v.log | show "bench $v"
done
if [ -f bournemouth-forced-aligner_b200/lib/libbfa_b200_prof.so ]; then
BFA_B200_LIB=$PWD/bournemouth-forced-aligner_b200/lib/libbfa_b200_prof.so timeout 300 python scripts/phase_prof.py 4096 > gpurun_out/phase.log 2>&1
cat gpurun_out/phase.log
fi
if [ $# -gt 0 ]; then "$@"; fi
