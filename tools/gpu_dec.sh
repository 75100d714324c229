#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_moses.py -m gpu -x -q -k "persistent_kernel or sample" > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/pytest_dec.log
timeout 200 python - > gpurun_out/dec_bench.log 2>&1 <<'PY'
import os, json, sys
sys.path.insert(0, os.getcwd())
import bench
for flag in ("2", "1", "0"):
    os.environ["MVAE_SAMPLE_PERSISTENT"] = flag
    r = bench.sampling_rate(n_total=0)
    print(flag, json.dumps({k: r[k] for k in ("value", "ms_per_batch", "mean_len")}), flush=True)
os.environ["MVAE_SAMPLE_PERSISTENT"] = "2"
r = bench.sampling_rate(batch=9472, n_total=0); print("9472", r["value"], flush=True)
r = bench.sampling_rate(batch=18944, n_total=0); print("18944", r["value"], flush=True)
PY
echo "bench rc $?"; cat gpurun_out/dec_bench.log | tail -8
