#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/rec_variants5.log; : > $L
run() { echo "== $1 debug=$2" >> $L; timeout 60 tools/$1 32 4096 120 1 $2 2>&1 | grep -E "variant 32|K-split vs pair" | head -3 >> $L; }
run rec_test 0
run rec_test $((2<<16))
cat $L
python bench.py --steps 20 --warmup 5 --samples 1000000 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
