#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/rec_trace.log; : > $L
for d in 0 64; do echo "== rec_test debug=$d" >> $L; timeout 60 tools/rec_test 32 4096 120 1 $d 2>&1 | grep -v "by t\|by pair\|by row\|by dG\|mismatch t" >> $L; done
cat $L
