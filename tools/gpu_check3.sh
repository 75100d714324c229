#!/bin/bash
# GPU-box pass: new tests, the bench line (with the L2 fabric probe), ncu --set full capture of one forward and one BPTT sweep.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cfgb.py -m gpu -q -k "graphed or tc_sgemm or phased" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_new.log
python bench.py --steps 20 --warmup 5 --samples 1000000 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_rec2_kernel --launch-skip 2 --launch-count 2 \
  -o gpurun_out/r02_rec2_full -f python tools/profile_step.py 4096 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
ls -la gpurun_out/*.ncu-rep
