"""Side measurement: the MOSES VAE step of bench.py alone.  python tools/bench_moses.py [mosesfile+head|mosesvae]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

for v in sys.argv[1:] or ["mosesfile+head", "mosesvae"]:
    print(json.dumps(bench.moses_step_rate(variant=v)), flush=True)
