"""CPU oracle for "Config A": the model exactly as shipped in the reference's models.py (TEST INFRASTRUCTURE ONLY).

numpy restatement (+ hand-derived backward) of
  models.py:126-135  MolEncoder.forward : Embedding(35,30) -> LSTM 3x72 -> ConvSELU(120->120,k18) -> ConvSELU(120->64,k18)
                                          -> ConvSELU(64->64,k18) -> Flatten -> Linear(1344,512)+SELU -> Lambda
  models.py:89-94    Lambda.forward     : mu, log_v = Linear(512,Z) x2 ; z = mu + exp(log_v/2) * (1e-2 * N(0,1))
  models.py:161-165  MolDecoder.forward : Linear(Z,Z)+SELU -> Repeat(120) -> LSTM 4x1024 -> Linear(1024,35) -> Softmax
  train.py:31-38     loss_function      : max_len*BCE(mean) + swapped KL
torch.nn.LSTM: gate rows i,f,g,o ; c' = f c + i g ; h' = o tanh(c')   (SURVEY.md A.1).
NB the conv "channels" are the 120 sequence positions and the conv length axis is the LSTM feature axis (72).
Pinned by tests/golden/make_golden_cfga.py (imports /root/reference/models.py) -> tests/golden/cfga_*.npz.
"""
from __future__ import annotations

import numpy as np

from .vae_oracle import (bce_grad_wrt_logits, bce_mean_times, conv1d_valid, conv1d_valid_bwd, kl_swapped, selu,
                         selu_grad, sigmoid, softmax_rows)


def cfga_shapes(T=120, Z=292, C=35, emb=30, eh=72, el=3, dh=1024, dl=4):
    s = {"encoder.embedding.weight": (C, emb)}
    for l in range(el):
        inp = emb if l == 0 else eh
        s[f"encoder.gru.weight_ih_l{l}"] = (4 * eh, inp)
        s[f"encoder.gru.weight_hh_l{l}"] = (4 * eh, eh)
        s[f"encoder.gru.bias_ih_l{l}"] = (4 * eh,)
        s[f"encoder.gru.bias_hh_l{l}"] = (4 * eh,)
    l1 = eh - 17; l2 = l1 - 17; l3 = l2 - 17
    s["encoder.conv_1.0.weight"] = (120, T, 18); s["encoder.conv_1.0.bias"] = (120,)
    s["encoder.conv_2.0.weight"] = (64, 120, 18); s["encoder.conv_2.0.bias"] = (64,)
    s["encoder.conv_3.0.weight"] = (64, 64, 18); s["encoder.conv_3.0.bias"] = (64,)
    s["encoder.dense_1.0.weight"] = (512, 64 * l3); s["encoder.dense_1.0.bias"] = (512,)
    s["encoder.lmbd.z_mean.weight"] = (Z, 512); s["encoder.lmbd.z_mean.bias"] = (Z,)
    s["encoder.lmbd.z_log_var.weight"] = (Z, 512); s["encoder.lmbd.z_log_var.bias"] = (Z,)
    s["decoder.latent_input.0.weight"] = (Z, Z); s["decoder.latent_input.0.bias"] = (Z,)
    for l in range(dl):
        inp = Z if l == 0 else dh
        s[f"decoder.gru.weight_ih_l{l}"] = (4 * dh, inp)
        s[f"decoder.gru.weight_hh_l{l}"] = (4 * dh, dh)
        s[f"decoder.gru.bias_ih_l{l}"] = (4 * dh,)
        s[f"decoder.gru.bias_hh_l{l}"] = (4 * dh,)
    s["decoder.decoded_mean.module.0.weight"] = (C, dh); s["decoder.decoded_mean.module.0.bias"] = (C,)
    return s


def make_cfga_params(seed, dtype=np.float32, **cfg):
    rng = np.random.Generator(np.random.PCG64(seed))
    shapes = cfga_shapes(**cfg)
    out = {}
    for k, shp in shapes.items():
        if k == "encoder.embedding.weight":
            out[k] = rng.standard_normal(shp).astype(dtype)
            continue
        if ".gru." in k:
            fan = shapes[k.rsplit(".", 1)[0] + ".weight_hh_l0"][1]
        elif k.endswith("weight"):
            fan = int(np.prod(shp[1:]))
        else:
            fan = int(np.prod(shapes[k.replace("bias", "weight")][1:]))
        b = 1.0 / np.sqrt(fan)
        out[k] = rng.uniform(-b, b, size=shp).astype(dtype)
    return out


def lstm_stack_forward(x0, T, Ws):
    """x0: (B,T,I) or (B,I) (time-invariant).  Ws: list of (w_ih, w_hh, b_ih, b_hh).  h0 = c0 = 0."""
    cache, inp = [], x0
    for (w_ih, w_hh, b_ih, b_hh) in Ws:
        H = w_hh.shape[1]
        B = inp.shape[0]
        dt = w_hh.dtype
        gi_all = inp @ w_ih.T + b_ih
        if inp.ndim == 2:
            gi_all = np.broadcast_to(gi_all[:, None, :], (B, T, 4 * H))
        h = np.zeros((B, H), dt); c = np.zeros((B, H), dt)
        hs = np.empty((B, T + 1, H), dt); hs[:, 0] = h
        sv = {k: np.empty((B, T, H), dt) for k in ("i", "f", "g", "o", "cprev", "tc")}
        for t in range(T):
            a = gi_all[:, t] + h @ w_hh.T + b_hh
            i, f, g, o = sigmoid(a[:, :H]), sigmoid(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]), sigmoid(a[:, 3 * H:])
            sv["cprev"][:, t] = c
            c = f * c + i * g
            tc = np.tanh(c)
            h = o * tc
            hs[:, t + 1] = h
            sv["i"][:, t], sv["f"][:, t], sv["g"][:, t], sv["o"][:, t], sv["tc"][:, t] = i, f, g, o, tc
        cache.append(dict(inp=inp, hs=hs, **sv))
        inp = hs[:, 1:]
    return inp, cache


def lstm_stack_backward(dout, Ws, cache):
    grads = [None] * len(Ws)
    for l in reversed(range(len(Ws))):
        w_ih, w_hh, b_ih, b_hh = Ws[l]
        c = cache[l]
        B, T, H = c["i"].shape
        dt = w_hh.dtype
        dG = np.empty((B, T, 4 * H), dt)
        dh = np.zeros((B, H), dt); dc = np.zeros((B, H), dt)
        for t in reversed(range(T)):
            dh = dh + dout[:, t]
            i, f, g, o, tc, cp = c["i"][:, t], c["f"][:, t], c["g"][:, t], c["o"][:, t], c["tc"][:, t], c["cprev"][:, t]
            do = dh * tc * o * (1 - o)
            dct = dc + dh * o * (1 - tc * tc)
            di = dct * g * i * (1 - i)
            df = dct * cp * f * (1 - f)
            dg = dct * i * (1 - g * g)
            dG[:, t] = np.concatenate([di, df, dg, do], 1)
            dc = dct * f
            dh = dG[:, t] @ w_hh
        dG2 = dG.reshape(B * T, 4 * H)
        g_ = {"w_hh": dG2.T @ c["hs"][:, :-1].reshape(B * T, H), "b_hh": dG2.sum(0), "b_ih": dG2.sum(0)}
        inp = c["inp"]
        if inp.ndim == 2:
            ds = dG.sum(1)
            g_["w_ih"] = ds.T @ inp
            dout = ds @ w_ih
        else:
            g_["w_ih"] = dG2.T @ inp.reshape(B * T, -1)
            dout = (dG2 @ w_ih).reshape(B, T, -1)
        grads[l] = g_
    return dout, grads


def cfga_step(P, ids, eps, max_len=120, eps_scale=1e-2, need_grads=True, el=3, dl=4):
    """ids (B,T) int; eps (B,Z) standard normal draws (models.py:92 scales them by 1e-2)."""
    dt = P["encoder.dense_1.0.weight"].dtype
    B, T = ids.shape
    C = P["decoder.decoded_mean.module.0.weight"].shape[0]
    E = P["encoder.embedding.weight"]
    emb = E[ids]
    We = [(P[f"encoder.gru.weight_ih_l{l}"], P[f"encoder.gru.weight_hh_l{l}"], P[f"encoder.gru.bias_ih_l{l}"],
           P[f"encoder.gru.bias_hh_l{l}"]) for l in range(el)]
    enc_out, ce = lstm_stack_forward(emb, T, We)                       # (B,T,72): conv channels = T positions
    a1, cols1 = conv1d_valid(enc_out, P["encoder.conv_1.0.weight"], P["encoder.conv_1.0.bias"]); h1 = selu(a1)
    a2, cols2 = conv1d_valid(h1, P["encoder.conv_2.0.weight"], P["encoder.conv_2.0.bias"]); h2 = selu(a2)
    a3, cols3 = conv1d_valid(h2, P["encoder.conv_3.0.weight"], P["encoder.conv_3.0.bias"]); h3 = selu(a3)
    flat = h3.reshape(B, -1)
    a4 = flat @ P["encoder.dense_1.0.weight"].T + P["encoder.dense_1.0.bias"]; h4 = selu(a4)
    mu = h4 @ P["encoder.lmbd.z_mean.weight"].T + P["encoder.lmbd.z_mean.bias"]
    lv = h4 @ P["encoder.lmbd.z_log_var.weight"].T + P["encoder.lmbd.z_log_var.bias"]
    std = np.exp(lv / 2.0)
    z = mu + std * (eps_scale * eps.astype(dt))
    a5 = z @ P["decoder.latent_input.0.weight"].T + P["decoder.latent_input.0.bias"]; zr = selu(a5)
    Wd = [(P[f"decoder.gru.weight_ih_l{l}"], P[f"decoder.gru.weight_hh_l{l}"], P[f"decoder.gru.bias_ih_l{l}"],
           P[f"decoder.gru.bias_hh_l{l}"]) for l in range(dl)]
    out, cd = lstm_stack_forward(zr, T, Wd)
    logits = out @ P["decoder.decoded_mean.module.0.weight"].T + P["decoder.decoded_mean.module.0.bias"]
    probs = softmax_rows(logits)
    onehot = np.zeros((B, T, C), dt); onehot[np.arange(B)[:, None], np.arange(T)[None, :], ids] = 1
    bce = bce_mean_times(probs, onehot, max_len)
    kl = kl_swapped(mu, lv)
    res = dict(probs=probs, mu=mu, logvar=lv, z=z, bce=float(bce), kl=float(kl), loss=float(bce + kl),
               argmax=probs.argmax(-1))
    if not need_grads:
        return res
    G = {}
    dl2 = bce_grad_wrt_logits(probs, onehot, max_len).astype(dt).reshape(B * T, C)
    G["decoder.decoded_mean.module.0.weight"] = dl2.T @ out.reshape(B * T, -1)
    G["decoder.decoded_mean.module.0.bias"] = dl2.sum(0)
    dout = (dl2 @ P["decoder.decoded_mean.module.0.weight"]).reshape(B, T, -1)
    dzr, gg = lstm_stack_backward(dout, Wd, cd)
    for l in range(dl):
        for nm, key in (("w_ih", "weight_ih"), ("w_hh", "weight_hh"), ("b_ih", "bias_ih"), ("b_hh", "bias_hh")):
            G[f"decoder.gru.{key}_l{l}"] = gg[l][nm]
    da5 = dzr * selu_grad(a5)
    G["decoder.latent_input.0.weight"] = da5.T @ z; G["decoder.latent_input.0.bias"] = da5.sum(0)
    dz = da5 @ P["decoder.latent_input.0.weight"]
    BZ = mu.size
    dmu = (dz + (-0.5 * (1.0 - np.exp(mu)) / BZ)).astype(dt)
    dlv = (lv / BZ + dz * (eps_scale * eps.astype(dt)) * std * 0.5).astype(dt)
    G["encoder.lmbd.z_mean.weight"] = dmu.T @ h4; G["encoder.lmbd.z_mean.bias"] = dmu.sum(0)
    G["encoder.lmbd.z_log_var.weight"] = dlv.T @ h4; G["encoder.lmbd.z_log_var.bias"] = dlv.sum(0)
    dh4 = dmu @ P["encoder.lmbd.z_mean.weight"] + dlv @ P["encoder.lmbd.z_log_var.weight"]
    da4 = dh4 * selu_grad(a4)
    G["encoder.dense_1.0.weight"] = da4.T @ flat; G["encoder.dense_1.0.bias"] = da4.sum(0)
    dh3 = (da4 @ P["encoder.dense_1.0.weight"]).reshape(h3.shape)
    da3 = dh3 * selu_grad(a3)
    dh2, G["encoder.conv_3.0.weight"], G["encoder.conv_3.0.bias"] = conv1d_valid_bwd(da3, cols3, P["encoder.conv_3.0.weight"], h2.shape)
    da2 = dh2 * selu_grad(a2)
    dh1, G["encoder.conv_2.0.weight"], G["encoder.conv_2.0.bias"] = conv1d_valid_bwd(da2, cols2, P["encoder.conv_2.0.weight"], h1.shape)
    da1 = dh1 * selu_grad(a1)
    denc, G["encoder.conv_1.0.weight"], G["encoder.conv_1.0.bias"] = conv1d_valid_bwd(da1, cols1, P["encoder.conv_1.0.weight"], enc_out.shape)
    demb, ge = lstm_stack_backward(denc, We, ce)
    for l in range(el):
        for nm, key in (("w_ih", "weight_ih"), ("w_hh", "weight_hh"), ("b_ih", "bias_ih"), ("b_hh", "bias_hh")):
            G[f"encoder.gru.{key}_l{l}"] = ge[l][nm]
    dE = np.zeros_like(E)
    np.add.at(dE, ids.reshape(-1), demb.reshape(B * T, -1))
    G["encoder.embedding.weight"] = dE
    res["grads"] = G
    return res
