#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/rec_variants2.log; : > $L
run() { echo "== $1 debug=$2" >> $L; timeout 60 tools/$1 32 4096 120 1 $2 2>&1 | grep -E "variant 32|K-split vs pair" | head -3 >> $L; }
run rec_test 0
run rec_test 4
run rec_test 32
run rec_test 36
run rec_test 1
run rec_test_o0n6 0
run rec_test_o0n6 $((3<<16))
run rec_test_o0n6 $((4<<16))
run rec_test_s2 0
cat $L
