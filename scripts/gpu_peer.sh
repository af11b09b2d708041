#!/bin/bash
# N-GPU: peer-mapped result arenas (--gather peer) against copy-engine pushes (--gather root), bench line and corpus driver
N=${1:-2}; CORPUS=${2:-262144}
mkdir -p gpurun_out
P=29611
timeout 300 python -m pytest tests/test_cuda_full.py -m gpu -x -q -k "caller_owned or pipelined" 2>&1 | tail -3
for g in peer root peer root; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-variants --gather $g > gpurun_out/bench_${g}_n$N.log 2> gpurun_out/bench_${g}_n$N.err; echo "bench $g N=$N rc=$?"
  python scripts/show_bench.py gpurun_out/bench_${g}_n$N.log; tail -2 gpurun_out/bench_${g}_n$N.err | cut -c1-300
  P=$((P+1))
done
for g in peer root; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --corpus $CORPUS --gather $g > gpurun_out/corpus_${g}_n$N.log 2> gpurun_out/corpus_${g}_n$N.err; echo "corpus $g N=$N rc=$?"
  tail -1 gpurun_out/corpus_${g}_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('n_gpus', 'align_ms_max_over_ranks', 'gather_tail_ms_after_last_kernel', 'wall_ms_incl_final_barrier', 'value', 'statuses_ok', 'gather_verified')})"
  tail -2 gpurun_out/corpus_${g}_n$N.err | cut -c1-300
  P=$((P+1))
done
