"""The reference's training-loop bodies, restated once so that the SAME loop code drives the reference modules (on torch
CPU, when tests/golden/make_golden_curves.py writes the fixtures) and the B200 drop-in modules (in tests/test_gpu_loss_curve.py).

  cfgb_loop   train.py:94-104     zero_grad; model(x); loss_function; backward; clip_grad_norm 3.0; Adam(8e-4).step
  moses_loop  moses_train_distrib_logp.py:289-338 (one outer iteration): the "aggressive encoder" step
              (kl_weight*kl + recon, clip 25, encoder Adam 8e-4) followed by the main step (recon only, clip 50, decoder
              Adam 5e-4), kl_weight following the script's recurrence min(kl_weight*0.1 + 1e-3, 1) (:302)
The normal draws are injected through `set_eps(step_index)` (the reference draws them with torch.randn_like; its stream
cannot be shared between devices), and the MOSES models run in eval() mode: the decoder's train-mode dropout mask stream of
cuDNN / ATen cannot be reproduced either, everything else of the loop is as shipped."""
import numpy as np
import torch

from oracle import moses_oracle as mo
from oracle import vae_oracle as vo

CFGB = dict(param_seed=701, B=32, Z=292, H=501, L=3, steps=30, lr=8e-4, clip=3.0, max_len=120)
MOSES = dict(param_seed=711, B=48, steps=15, lr_enc=8e-4, lr_dec=5e-4, clip_agg=25.0, clip_main=50.0, kl0=0.1)


def cfgb_data(step):
    ids, onehot, eps = vo.make_batch(7100 + step, CFGB["B"], latent=CFGB["Z"], dtype=np.float32)
    return ids, onehot, eps


def cfgb_loop(model, loss_function, make_input, set_eps, zero_grad=None, clip_and_step=None, steps=None):
    """Returns the per-step losses.  Default optimiser: torch.optim.Adam(8e-4) + clip_grad_norm_(3.0) as train.py:81,102-104;
    zero_grad / clip_and_step replace them (the fused clip + Adam step of molecular-vae_b200.optim.FusedOptimizer)."""
    c = CFGB
    if clip_and_step is None:
        optimizer = torch.optim.Adam(model.parameters(), lr=c["lr"])
        zero_grad = optimizer.zero_grad

        def clip_and_step():
            torch.nn.utils.clip_grad_norm_(model.parameters(), c["clip"])
            optimizer.step()
    losses = []
    for step in range(steps or c["steps"]):
        ids, onehot, eps = cfgb_data(step)
        x = make_input(onehot)
        set_eps(eps)
        zero_grad()
        recon, mu, logvar = model(x)
        loss = loss_function(recon, x, mu, logvar)
        loss.backward()
        clip_and_step()
        losses.append(float(loss.detach()))
    return np.array(losses)


def moses_data(step):
    seqs, eps, pad = mo.make_moses_batch(7200 + step, MOSES["B"], dtype=np.float32)
    return seqs, eps, pad


def moses_loop(model, to_dev, set_eps, steps=None):
    """Returns (aggressive-step losses, main-step recon losses, kl values)."""
    c = MOSES
    enc = torch.optim.Adam(model.encoder.parameters(), lr=c["lr_enc"])
    dec = torch.optim.Adam(model.decoder.parameters(), lr=c["lr_dec"])
    klw = c["kl0"]
    agg, main, kls = [], [], []
    for step in range(steps or c["steps"]):
        seqs, eps, pad = moses_data(step)
        x = [to_dev(torch.from_numpy(s)) for s in seqs]
        # aggressive encoder step (:292-310)
        enc.zero_grad(); dec.zero_grad()
        set_eps(eps)
        out = model(x)
        kl, recon = out[0], out[1]
        klw = min(klw * 1e-1 + 1e-3, 1)
        loss = klw * kl + recon
        loss.backward()
        torch.nn.utils.clip_grad_norm_((p for p in model.parameters() if p.requires_grad), c["clip_agg"])
        enc.step()
        agg.append(float(loss.detach())); kls.append(float(kl.detach()))
        # main step (:321-338, epoch < 10: decoder only)
        enc.zero_grad(); dec.zero_grad()
        set_eps(eps)
        out = model(x)
        recon = out[1]
        recon.backward()
        torch.nn.utils.clip_grad_norm_((p for p in model.parameters() if p.requires_grad), c["clip_main"])
        dec.step()
        main.append(float(recon.detach()))
    return np.array(agg), np.array(main), np.array(kls)
