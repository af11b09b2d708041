#!/bin/bash
# A/B: device-timed bench of the default library and of every lib/libbfa_b200_<variant>.so given as argument
mkdir -p gpurun_out
run() {
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$1: value %.3e ms/step %.4f frac %.3f isolated %.4f fill %.4f host_us %.1f plain %.4f' % (d['value'], d['ms_per_step'], r['frac'], r.get('isolated_launch_ms', 0), r.get('fill_phase', {}).get('kernel_ms', 0), d.get('host_enqueue_us_per_step', 0), d['uninstrumented_step']['ms_per_step']))"
}
run default
for v in "$@"; do
  BFA_B200_LIB=$PWD/bournemouth-forced-aligner_b200/lib/libbfa_b200_$v.so run $v
done
run default
