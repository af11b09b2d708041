#!/bin/bash
# round 2 evidence run: parity tests, full bench line (variants, e2e, cpu baseline), reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_full.log
tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
tail -1 gpurun_out/bench_ref.log | cut -c1-1500
if [ $# -gt 0 ]; then "$@"; fi
