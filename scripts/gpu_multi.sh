#!/bin/bash
# N-GPU runs: the contract bench line and the BASELINE config 5 corpus driver
N=${1:-2}; CORPUS=${2:-262144}
mkdir -p gpurun_out
P=29511
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
python scripts/show_bench.py gpurun_out/bench_n$N.log
tail -3 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --corpus $CORPUS > gpurun_out/corpus_n$N.log 2> gpurun_out/corpus_n$N.err; echo "corpus N=$N rc=$?"
tail -1 gpurun_out/corpus_n$N.log | cut -c1-1200
tail -3 gpurun_out/corpus_n$N.err
