#!/bin/bash
# 2-GPU pass: the data-parallel bench line with the single-graph step and with the per-phase graphs.
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --samples 200000; }
MVAE_DP_GRAPH=1 run 29511 > gpurun_out/bench_n2_graph.json 2> gpurun_out/bench_n2_graph.err; echo "n2 graph rc $?"
MVAE_DP_GRAPH=0 run 29512 > gpurun_out/bench_n2_phased.json 2> gpurun_out/bench_n2_phased.err; echo "n2 phased rc $?"
tail -c 600 gpurun_out/bench_n2_graph.err
