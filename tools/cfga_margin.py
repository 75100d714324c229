"""Gradient error margins of Config A's bf16 step against the float64 oracle (same cases as
tests/test_gpu_cfga.py::test_fused_step_bf16_full_config): prints the five largest relative L2 errors per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cfga_oracle as ca
from tests.util_gpu import grads_of, load_pkg, rel_l2
from tests.test_gpu_cfga import _build, _case, _fused

m = load_pkg()
for B in (64, 136):
    P, P64, ids, onehot, eps = _case(71, 72 + B, B, 292, 72, 3, 1024, 4)
    ref = ca.cfga_step(P64, ids.astype(np.int64), eps.astype(np.float64))
    model = _build(m, P, 292, 72, 3, 1024, 4, "bf16")
    sc = _fused(model, ids, eps)
    errs = sorted(((rel_l2(g, ref["grads"][k]), k) for k, g in grads_of(model).items()), reverse=True)
    print("B", B, "mode", os.environ.get("MVAE_LSTM_GATE_X8", "default"), "loss rel", abs(sc[0] - ref["loss"]) / abs(ref["loss"]),
          " ".join(f"{k}={e:.5f}" for e, k in errs[:5]), flush=True)
