#!/bin/bash
# Final evidence pass of the round: parity suite, bench line, ncu launch lists (Config B, MOSES), ncu --set full of the sweeps.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python tools/profile_step.py 4096 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_moses.csv \
  python tools/profile_moses.py 4096 2 > gpurun_out/ncu_launch_moses.log 2>&1; echo "ncu moses rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_sampler.csv \
  python tools/profile_sampler.py 8192 100 > gpurun_out/ncu_launch_sampler.log 2>&1; echo "ncu sampler rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_rec2_kernel --launch-skip 2 --launch-count 2 \
  -o gpurun_out/r02_rec2_full -f python tools/profile_step.py 4096 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
python tools/ncu_summary.py full gpurun_out/r02_rec2_full.ncu-rep > gpurun_out/r02_rec2_ncu_full.txt 2>&1
python tools/ncu_stalls.py gpurun_out/r02_rec2_full.ncu-rep molecular-vae_b200/_build/gru_rec2.o 22 > gpurun_out/r02_rec2_stalls.txt 2>&1
