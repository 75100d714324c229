"""Side measurement for the models.py path ("Config A": LSTM 3x72 encoder, LSTM 4x1024 decoder) on ONE GPU; not the
headline metric.  Fused ELBO step (fwd + loss + bwd) at the given batch, CUDA-graph replay, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import molecular_vae_b200 as m
from oracle import vae_oracle as vo

GFLOP_PER_MOL = 21.3859


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    torch.manual_seed(42)
    model = m.models.MolecularVAE(precision=prec).cuda()
    ids, _, eps = vo.make_batch(1, B)
    ids, eps = torch.from_numpy(ids).cuda(), torch.from_numpy(eps).cuda()
    out = model.elbo_step(ids, eps, max_len=120); torch.cuda.synchronize(); model.engine(B).check_device_error()
    print("scalars", out.cpu().numpy(), "graph nodes", m._lib.lib.mvae_graph_num_kernel_nodes(model.engine(B)._graph), flush=True)
    model.elbo_step(ids, eps, max_len=120); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        model.engine(B).launch_graph()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    mols = B / ms * 1e3
    print(f"cfga train step {prec} B={B}: {ms:.2f} ms/step {mols:.0f} molecules/s  "
          f"{mols * GFLOP_PER_MOL / 1e3:.1f} TFLOP/s algorithmic  ws={model.engine(B).ws_bytes / 1e9:.1f} GB", flush=True)


if __name__ == "__main__":
    main()
