"""Error budget of Config A's bf16 mode (TEST INFRASTRUCTURE, CPU only; not collected by pytest).

Restates oracle/cfga_oracle.py::cfga_step in float64 with a bf16 rounding hook at every place where cfga.cu stores a tensor
as bf16 (operands of the tcgen05 GEMMs, saved gates, gradients handed between layers), so that the contribution of each
group of roundings to the gradient error of the encoder LSTM -- the tensors that use 0.90-0.95 of the 1e-2 budget on the
GPU (tools/cfga_margin.py) -- can be measured by switching groups on and off.  fp32 accumulation is modelled as exact.

    python -m tests.cfga_bf16_budget [B]          # prints one line per experiment

Rounding sites (cfga.cu):  e_* encoder LSTM, c_* convolutions, d_* decoder LSTM + head
  *_w   weights as bf16 operands            *_gi  input projections stored bf16 (not decoder layer 0: fp32)
  *_h   hidden state stored bf16            *_sv  saved gates (i, f, g, o, c_prev, tanh c) bf16
  *_dG  pre-activation gradients bf16       *_dX  gradient handed to the layer below bf16
  c_cols im2col operand bf16   c_da  conv pre-activation gradient bf16   c_dcols  da W as bf16   d_dlog  d(logits) bf16
"""
import sys

import numpy as np

from oracle import cfga_oracle as ca
from oracle.vae_oracle import (bce_grad_wrt_logits, kl_swapped, make_batch, selu, selu_grad, sigmoid, softmax_rows)

SITES = ["e_w", "e_gi", "e_h", "e_sv", "e_dG", "e_dX", "c_w", "c_cols", "c_da", "c_dcols",
         "d_w", "d_gi", "d_h", "d_sv", "d_dlog", "d_dG", "d_dX"]
GROUPS = {"encoder forward (e_w e_gi e_h)": ["e_w", "e_gi", "e_h"], "encoder saved gates (e_sv)": ["e_sv"],
          "encoder BPTT (e_dG e_dX)": ["e_dG", "e_dX"], "conv forward (c_w c_cols)": ["c_w", "c_cols"],
          "conv backward (c_da c_dcols)": ["c_da", "c_dcols"], "decoder forward (d_w d_gi d_h)": ["d_w", "d_gi", "d_h"],
          "decoder saved gates (d_sv)": ["d_sv"], "decoder BPTT + head (d_dlog d_dG d_dX)": ["d_dlog", "d_dG", "d_dX"]}


def bf16(x):
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32)
    u = (u + (((u >> 16) & 1) + 0x7FFF)) & 0xFFFF0000
    return u.view(np.float32).astype(np.float64)


class Rounder:
    """coded_sv: the saved activations keep the relative precision of their complements -- a sigmoid s > 0.5 is stored as its
    complement 1 - s, a tanh value |g| > 0.5 as 1 - |g| -- so that s (1 - s) and 1 - g^2 do not lose their leading digits
    when the activation saturates."""

    def __init__(self, active, coded_sv=False):
        self.active = set(active)
        self.coded_sv = coded_sv

    def __call__(self, site, x):
        return bf16(x) if site in self.active else x

    def sv(self, site, kind, x):
        if site not in self.active:
            return x
        if not self.coded_sv or kind == "lin":
            return bf16(x)
        if kind == "sig":
            return np.where(x <= 0.5, bf16(x), 1.0 - bf16(1.0 - x))
        return np.where(np.abs(x) <= 0.5, bf16(x), np.sign(x) * (1.0 - bf16(1.0 - np.abs(x))))


def lstm_fwd(r, pre, x0, T, Ws, gi0_exact):
    """x0 (B,T,I) already as stored, or (B,I) time-invariant.  gi0_exact: layer 0's projection uses the fp32 parameters."""
    cache, inp = [], x0
    for l, (w_ih, w_hh, b_ih, b_hh) in enumerate(Ws):
        H = w_hh.shape[1]
        B = inp.shape[0]
        whh = r(pre + "w", w_hh)
        if l == 0 and gi0_exact:
            gi = inp @ w_ih.T + (b_ih + b_hh)
            if inp.ndim == 2:
                gi = np.broadcast_to(gi[:, None, :], (B, T, 4 * H))      # decoder: fp32 gi0, not rounded
            else:
                gi = r(pre + "gi", gi)                                    # encoder: table gathered into bf16
        else:
            gi = r(pre + "gi", inp @ r(pre + "w", w_ih).T + (b_ih + b_hh))
        h = np.zeros((B, H)); c = np.zeros((B, H))
        hs = np.empty((B, T + 1, H)); hs[:, 0] = h
        sv = {k: np.empty((B, T, H)) for k in ("i", "f", "g", "o", "cprev", "tc")}
        for t in range(T):
            a = gi[:, t] + h @ whh.T
            i, f, g, o = sigmoid(a[:, :H]), sigmoid(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]), sigmoid(a[:, 3 * H:])
            sv["cprev"][:, t] = r.sv(pre + "sv", "lin", c)
            c = f * c + i * g
            tc = np.tanh(c)
            h = r(pre + "h", o * tc)
            hs[:, t + 1] = h
            for k, v in (("i", i), ("f", f), ("g", g), ("o", o), ("tc", tc)):
                sv[k][:, t] = r.sv(pre + "sv", "tanh" if k in ("g", "tc") else "sig", v)
        cache.append(dict(inp=inp, hs=hs, **sv))
        inp = hs[:, 1:]
    return inp, cache


def lstm_bwd(r, pre, dout, Ws, cache):
    """dout as stored (already rounded by the caller's *_dX site)."""
    grads = [None] * len(Ws)
    for l in reversed(range(len(Ws))):
        w_ih, w_hh, b_ih, b_hh = Ws[l]
        whh = r(pre + "w", w_hh)
        c = cache[l]
        B, T, H = c["i"].shape
        dG = np.empty((B, T, 4 * H))
        dh = np.zeros((B, H)); dc = np.zeros((B, H))
        for t in reversed(range(T)):
            dh = dh + dout[:, t]
            i, f, g, o, tc, cp = c["i"][:, t], c["f"][:, t], c["g"][:, t], c["o"][:, t], c["tc"][:, t], c["cprev"][:, t]
            do = dh * tc * o * (1 - o)
            dct = dc + dh * o * (1 - tc * tc)
            dG[:, t] = r(pre + "dG", np.concatenate([dct * g * i * (1 - i), dct * cp * f * (1 - f), dct * i * (1 - g * g), do], 1))
            dc = dct * f
            dh = dG[:, t] @ whh
        dG2 = dG.reshape(B * T, 4 * H)
        g_ = {"w_hh": dG2.T @ c["hs"][:, :-1].reshape(B * T, H), "b_hh": dG2.sum(0), "b_ih": dG2.sum(0)}
        inp = c["inp"]
        if inp.ndim == 2:                       # decoder layer 0: fp32 time sum, fp32-class products with the fp32 weights
            ds = dG.sum(1)
            g_["w_ih"] = ds.T @ inp
            dout = ds @ w_ih
        elif l == 0:                            # encoder layer 0: table gradient, fp32-class products with the fp32 weights
            g_["w_ih"] = dG2.T @ inp.reshape(B * T, -1)
            dout = (dG2 @ w_ih).reshape(B, T, -1)
        else:
            g_["w_ih"] = dG2.T @ inp.reshape(B * T, -1)
            dout = r(pre + "dX", (dG2 @ r(pre + "w", w_ih)).reshape(B, T, -1))
        grads[l] = g_
    return dout, grads


def conv_fwd(r, x, w, b):
    B, Cin, L = x.shape
    Cout, _, K = w.shape
    Lo = L - K + 1
    idx = np.arange(Lo)[:, None] + np.arange(K)[None, :]
    cols = r("c_cols", x[:, :, idx].transpose(0, 2, 1, 3).reshape(B * Lo, Cin * K))
    y = cols @ r("c_w", w.reshape(Cout, Cin * K)).T + b
    return y.reshape(B, Lo, Cout).transpose(0, 2, 1), cols


def conv_bwd(r, da, cols, w, x_shape):
    B, Cin, L = x_shape
    Cout, _, K = w.shape
    Lo = L - K + 1
    da2 = r("c_da", da.transpose(0, 2, 1).reshape(B * Lo, Cout))
    dw = (da2.T @ cols).reshape(w.shape)
    db = da2.sum(0)
    dcols = r("c_dcols", da2 @ r("c_w", w.reshape(Cout, Cin * K))).reshape(B, Lo, Cin, K)
    dx = np.zeros(x_shape)
    for k in range(K):
        dx[:, :, k:k + Lo] += dcols[:, :, :, k].transpose(0, 2, 1)
    return dx, dw, db


def step(P, ids, eps, active, max_len=120, eps_scale=1e-2, el=3, dl=4, coded_sv=False):
    r = Rounder(active, coded_sv)
    B, T = ids.shape
    C = P["decoder.decoded_mean.module.0.weight"].shape[0]
    E = P["encoder.embedding.weight"]
    emb = E[ids]
    We = [(P[f"encoder.gru.weight_ih_l{l}"], P[f"encoder.gru.weight_hh_l{l}"], P[f"encoder.gru.bias_ih_l{l}"],
           P[f"encoder.gru.bias_hh_l{l}"]) for l in range(el)]
    enc_out, ce = lstm_fwd(r, "e_", emb, T, We, True)
    a1, cols1 = conv_fwd(r, enc_out, P["encoder.conv_1.0.weight"], P["encoder.conv_1.0.bias"]); h1 = selu(a1)
    a2, cols2 = conv_fwd(r, h1, P["encoder.conv_2.0.weight"], P["encoder.conv_2.0.bias"]); h2 = selu(a2)
    a3, cols3 = conv_fwd(r, h2, P["encoder.conv_3.0.weight"], P["encoder.conv_3.0.bias"]); h3 = selu(a3)
    flat = h3.reshape(B, -1)
    a4 = flat @ P["encoder.dense_1.0.weight"].T + P["encoder.dense_1.0.bias"]; h4 = selu(a4)
    mu = h4 @ P["encoder.lmbd.z_mean.weight"].T + P["encoder.lmbd.z_mean.bias"]
    lv = h4 @ P["encoder.lmbd.z_log_var.weight"].T + P["encoder.lmbd.z_log_var.bias"]
    std = np.exp(lv / 2.0)
    z = mu + std * (eps_scale * eps)
    a5 = z @ P["decoder.latent_input.0.weight"].T + P["decoder.latent_input.0.bias"]; zr = selu(a5)
    Wd = [(P[f"decoder.gru.weight_ih_l{l}"], P[f"decoder.gru.weight_hh_l{l}"], P[f"decoder.gru.bias_ih_l{l}"],
           P[f"decoder.gru.bias_hh_l{l}"]) for l in range(dl)]
    out, cd = lstm_fwd(r, "d_", zr, T, Wd, True)
    wfc = r("d_w", P["decoder.decoded_mean.module.0.weight"])
    probs = softmax_rows(out @ wfc.T + P["decoder.decoded_mean.module.0.bias"])
    onehot = np.zeros((B, T, C)); onehot[np.arange(B)[:, None], np.arange(T)[None, :], ids] = 1
    G = {}
    dl2 = r("d_dlog", bce_grad_wrt_logits(probs, onehot, max_len).reshape(B * T, C))
    G["decoder.decoded_mean.module.0.weight"] = dl2.T @ out.reshape(B * T, -1)
    dout = r("d_dX", (dl2 @ wfc).reshape(B, T, -1))
    dzr, gg = lstm_bwd(r, "d_", dout, Wd, cd)
    for l in range(dl):
        G[f"decoder.gru.weight_hh_l{l}"] = gg[l]["w_hh"]
    G["decoder.gru.weight_ih_l0"] = gg[0]["w_ih"]
    da5 = dzr * selu_grad(a5)
    G["decoder.latent_input.0.weight"] = da5.T @ z
    dz = da5 @ P["decoder.latent_input.0.weight"]
    BZ = mu.size
    dmu = dz + (-0.5 * (1.0 - np.exp(mu)) / BZ)
    dlv = lv / BZ + dz * (eps_scale * eps) * std * 0.5
    G["encoder.lmbd.z_mean.weight"] = dmu.T @ h4
    dh4 = dmu @ P["encoder.lmbd.z_mean.weight"] + dlv @ P["encoder.lmbd.z_log_var.weight"]
    da4 = dh4 * selu_grad(a4)
    G["encoder.dense_1.0.weight"] = da4.T @ flat
    dh3 = (da4 @ P["encoder.dense_1.0.weight"]).reshape(h3.shape)
    dh2, G["encoder.conv_3.0.weight"], _ = conv_bwd(r, dh3 * selu_grad(a3), cols3, P["encoder.conv_3.0.weight"], h2.shape)
    dh1, G["encoder.conv_2.0.weight"], _ = conv_bwd(r, dh2 * selu_grad(a2), cols2, P["encoder.conv_2.0.weight"], h1.shape)
    denc, G["encoder.conv_1.0.weight"], _ = conv_bwd(r, dh1 * selu_grad(a1), cols1, P["encoder.conv_1.0.weight"], enc_out.shape)
    G["_denc"] = denc
    demb, ge = lstm_bwd(r, "e_", r("e_dX", denc), We, ce)
    for l in range(el):
        G[f"encoder.gru.weight_hh_l{l}"] = ge[l]["w_hh"]
    dE = np.zeros_like(E)
    np.add.at(dE, ids.reshape(-1), demb.reshape(B * T, -1))
    G["encoder.embedding.weight"] = dE
    return G


KEYS = ["decoder.gru.weight_hh_l3", "decoder.gru.weight_hh_l0", "decoder.latent_input.0.weight", "encoder.lmbd.z_mean.weight",
        "encoder.dense_1.0.weight", "encoder.conv_3.0.weight", "encoder.conv_1.0.weight", "_denc", "encoder.gru.weight_hh_l2",
        "encoder.gru.weight_hh_l0", "encoder.embedding.weight"]


def rel(a, b):
    return float(np.sqrt(((a - b) ** 2).sum()) / np.sqrt((b ** 2).sum()))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    P = {k: v.astype(np.float64) for k, v in ca.make_cfga_params(71, dtype=np.float32).items()}
    ids, _, eps = make_batch(72 + B, B, latent=292, dtype=np.float32)
    ids, eps = ids.astype(np.int64), eps.astype(np.float64)
    exact = step(P, ids, eps, [])
    ref = ca.cfga_step(P, ids, eps)["grads"]
    print("restatement vs oracle:", max(rel(exact[k], ref[k]) for k in KEYS if k in ref), flush=True)
    print("columns:", " ".join(k.replace("encoder.", "e.").replace("decoder.", "d.").replace("weight", "w") for k in KEYS), flush=True)

    def run(name, active, coded_sv=False):
        G = step(P, ids, eps, active, coded_sv=coded_sv)
        print(f"{name:58s}", " ".join(f"{rel(G[k], exact[k]):.5f}" for k in KEYS), flush=True)

    run("all sites (the bf16 mode)", SITES)
    run("all sites, saved gates complement-coded", SITES, True)
    if len(sys.argv) > 2 and sys.argv[2] == "short":
        return
    for name, g in GROUPS.items():
        run("only   " + name, g)
    for name, g in GROUPS.items():
        run("all but " + name, [s for s in SITES if s not in g])


if __name__ == "__main__":
    main()
