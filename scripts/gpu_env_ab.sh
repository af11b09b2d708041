#!/bin/bash
# A/B over an environment variable: gpu_env_ab.sh VAR v1 v2 ...
var=$1; shift
for v in "$@" "$1"; do
  env $var=$v python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$var=$v: value %.3e ms/step %.4f kernel_ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
