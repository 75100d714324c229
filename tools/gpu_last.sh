#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_cfga.py -m gpu -q > gpurun_out/pytest_cfga.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_cfga.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('cfga_step'), d['sampling']['value'], d['moses_step']['value'])
PY
