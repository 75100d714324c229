#!/bin/bash
# GPU-box pass: parity tests, bench line, ncu launch lists of the Config-B and MOSES steps, Config-A side measurement.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --samples 1000000 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
python tools/bench_cfga.py bf16 4096 3 > gpurun_out/cfga.log 2>&1; echo "cfga rc $?"; tail -1 gpurun_out/cfga.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python tools/profile_step.py 4096 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_moses.csv \
  python tools/profile_moses.py 4096 2 > gpurun_out/ncu_launch_moses.log 2>&1; echo "ncu moses rc $?"
