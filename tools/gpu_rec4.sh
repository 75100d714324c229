#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/rec_variants3.log; : > $L
echo "== rec_test (st.async exchange, publisher: counter before tempty)" >> $L; timeout 60 tools/rec_test 32 4096 120 1 0 2>&1 | grep -E "variant 32|K-split vs pair" >> $L
cat $L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --samples 1000000 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
