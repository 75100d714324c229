#!/bin/bash
# 8-GPU pass: the data-parallel bench line (single-graph step), bounded.
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --samples 800000 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 rc $?"
tail -c 300 gpurun_out/bench_n8.err
