#!/bin/bash
# A/B: device-timed bench of the default library and of every lib/libbfa_b200_<variant>.so given as argument
mkdir -p gpurun_out
run() {
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1: value %.3e ms/step %.4f kernel_ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
}
run default
for v in "$@"; do
  BFA_B200_LIB=$PWD/bournemouth-forced-aligner_b200/lib/libbfa_b200_$v.so run $v
done
run default
