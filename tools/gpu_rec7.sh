#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/rec_variants6.log; : > $L
run() { echo "== $1 debug=$2 box=$3" >> $L; timeout 60 tools/$1 32 4096 120 1 $2 $3 2>&1 | grep -E "variant 32|K-split vs pair" | head -3 >> $L; }
run rec_test_v1 0 0
run rec_test 0 0
run rec_test 0 64
run rec_test 0 32
cat $L
