#!/bin/bash
# 2-GPU pass: the data-parallel bench line (single-graph step); bounded so that a hang costs minutes, not the call limit.
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --samples 200000 > gpurun_out/bench_n2_graph.json 2> gpurun_out/bench_n2_graph.err; echo "n2 graph rc $?"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "n2 ref rc $?"
tail -c 300 gpurun_out/bench_n2_graph.err
