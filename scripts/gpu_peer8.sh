#!/bin/bash
# N-GPU evidence for the peer-mapped result arenas: bench line (peer, root) and the config-5 corpus (peer)
N=${1:-8}; CORPUS=${2:-1048576}
mkdir -p gpurun_out
P=29711
for g in peer root; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-variants --gather $g > gpurun_out/bench_${g}_n$N.log 2> gpurun_out/bench_${g}_n$N.err; echo "bench $g N=$N rc=$?"
  python scripts/show_bench.py gpurun_out/bench_${g}_n$N.log; tail -2 gpurun_out/bench_${g}_n$N.err | cut -c1-300
  P=$((P+1))
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --corpus $CORPUS --gather peer > gpurun_out/corpus_peer_n$N.log 2> gpurun_out/corpus_peer_n$N.err; echo "corpus peer N=$N rc=$?"
tail -1 gpurun_out/corpus_peer_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('n_gpus', 'align_ms_max_over_ranks', 'gather_tail_ms_after_last_kernel', 'wall_ms_incl_final_barrier', 'value', 'statuses_ok', 'gather_verified')})"
tail -2 gpurun_out/corpus_peer_n$N.err | cut -c1-300
