#!/bin/bash
# recurrence-kernel timing variants (standalone tool) + the parity suite with the default build
mkdir -p gpurun_out
for b in rec_test_base rec_test rec_test_nst4 rec_test_f3; do
  echo "== $b" >> gpurun_out/rec_variants.log
  timeout 120 tools/$b 32 4096 120 1 2>&1 | grep -E "variant 32|K-split vs pair|mismatch t" | head -8 >> gpurun_out/rec_variants.log
done
cat gpurun_out/rec_variants.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --samples 1000000 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
