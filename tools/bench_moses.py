"""Side measurements for the MOSES VAE path (BASELINE.json configs[3] / configs[4] shapes on ONE GPU; not the headline
metric): fused train step at batch 4096 and greedy / multinomial sampling at batch 8192 x 100 steps (hugesample.py:113).
Each workload is captured into a torch CUDA graph and replayed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import molecular_vae_b200 as m
from oracle import moses_oracle as mo
from tests.test_gpu_moses import _Vocab


def timed(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    torch.manual_seed(0)
    model = m.mosesvae.VAE(_Vocab(), precision=prec).cuda()
    seqs, eps, pad = mo.make_moses_batch(1, B)
    x = [torch.from_numpy(s).cuda() for s in seqs]
    eps = torch.from_numpy(eps).cuda()
    tokens = sum(len(s) for s in seqs)
    model.elbo_step(x, kl_weight=0.1, eps=eps); torch.cuda.synchronize(); model.check_device_error()
    _, ids_p, lens_p = model._pack(x)
    params = model.ordered_params()
    P, G = [p.data for p in params], [p.grad for p in params]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        model._run(P, G, ids_p, lens_p, eps, 0.1, 1.0, False)
        out = model._last_scalars
    g.replay(); torch.cuda.synchronize()
    ms = timed(g.replay, 5)
    print(f"moses train step {prec} B={B} T={max(len(s) for s in seqs)} mean_len={tokens / B:.1f}: {ms:.2f} ms/step "
          f"{B / ms * 1e3:.0f} molecules/s  scalars={out.cpu().numpy()}", flush=True)
    SB = 8192
    z = torch.randn(SB, 160, device="cuda")
    for greedy in (True, False):
        model.sample_ids(SB, max_len=100, z=z, greedy=greedy, seed=3); torch.cuda.synchronize(); model.check_device_error()
        gs = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gs):
            ids, lens, _ = model.sample_ids(SB, max_len=100, z=z, greedy=greedy, seed=3, use_graph=False)
        gs.replay(); torch.cuda.synchronize()
        ms = timed(gs.replay, 3)
        print(f"moses sample {prec} {'greedy' if greedy else 'multinomial'} B={SB} max_len=100: {ms:.2f} ms "
              f"{SB / ms * 1e3:.0f} SMILES/s  mean_len={lens.float().mean().item():.1f}", flush=True)


if __name__ == "__main__":
    main()
