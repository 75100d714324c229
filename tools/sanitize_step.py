"""Tiny fused step for compute-sanitizer (memcheck): full Config B shapes at batch 8, bf16 and fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import molecular_vae_b200 as m
from oracle import vae_oracle as vo
for prec in ("bf16", "fp32"):
    model = m.VAE(latent=292, precision=prec).cuda()
    ids, _, eps = vo.make_batch(5, 8)
    out = model.elbo_step(torch.from_numpy(ids).cuda(), torch.from_numpy(eps).cuda(), use_graph=False)
    torch.cuda.synchronize()
    model.engine(8).check_device_error()
    print(prec, out.cpu().numpy())
    z = torch.randn(8, 292, device="cuda")
    print(model.decode_greedy(z)[0, :10].cpu().numpy())
