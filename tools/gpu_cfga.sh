#!/bin/bash
# Config A check: parity margins and step rate with the LSTM cell kernel variants (MVAE_LSTM_GATE_X8 = 0 / 1 / 2 / 3)
mkdir -p gpurun_out
for mode in 0 1 2 3; do MVAE_LSTM_GATE_X8=$mode timeout 200 python tools/cfga_margin.py 2>&1 | grep "^B "; done > gpurun_out/cfga_margin.log
cat gpurun_out/cfga_margin.log | cut -c1-330
timeout 300 python -m pytest tests/test_gpu_cfga.py -m gpu -q > gpurun_out/pytest_cfga.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_cfga.log
( for mode in 1 2; do MVAE_LSTM_GATE_X8=$mode timeout 120 python tools/bench_cfga.py bf16 4096 3 | tail -1 | sed "s/^/gate_x8=$mode: /"; done ) > gpurun_out/cfga_bench.log 2>&1
cat gpurun_out/cfga_bench.log
