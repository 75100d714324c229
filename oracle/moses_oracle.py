"""CPU oracle for the MOSES-style character VAE step (TEST INFRASTRUCTURE ONLY; see oracle/vae_oracle.py header).

numpy restatement of the reference's mosesvae.py forward (+ hand-derived backward):
  mosesvae.py:142-164  forward_encoder : emb -> packed GRU(V->256) -> last h -> MLP mu / MLP logvar -> z, KL
  mosesvae.py:166-199  forward_decoder : pad, emb, cat z, h0 = decoder_lat(z) x3, GRU 3x512, fc, shifted CE (ignore pad)
Packed-sequence semantics (torch pack_sequence / pack_padded_sequence): sequence b only advances for t < L_b, its
final hidden is the state after L_b steps, padded outputs are zero.  Dropout is the identity here (parity runs use
eval mode; the train-mode mask is an injected input of the CUDA path).

Parity pin: tests/golden/make_golden_moses.py imports /root/reference/mosesvae.py + vocab.py and writes
tests/golden/moses_*.npz; tests/test_oracle_golden.py holds this file to those fixtures.
"""
from __future__ import annotations

import numpy as np

from .vae_oracle import sigmoid


def moses_shapes(V=34, d_z=160, q_h=256, d_h=512, d_layers=3, mlp=256):
    s = {"x_emb.weight": (V, V),
         "encoder_rnn.weight_ih_l0": (3 * q_h, V), "encoder_rnn.weight_hh_l0": (3 * q_h, q_h),
         "encoder_rnn.bias_ih_l0": (3 * q_h,), "encoder_rnn.bias_hh_l0": (3 * q_h,),
         "q_mu.0.weight": (mlp, q_h), "q_mu.0.bias": (mlp,), "q_mu.2.weight": (d_z, mlp), "q_mu.2.bias": (d_z,),
         "q_logvar.0.weight": (mlp, q_h), "q_logvar.0.bias": (mlp,), "q_logvar.2.weight": (d_z, mlp),
         "q_logvar.2.bias": (d_z,)}
    for l in range(d_layers):
        inp = V + d_z if l == 0 else d_h
        s[f"decoder_rnn.weight_ih_l{l}"] = (3 * d_h, inp)
        s[f"decoder_rnn.weight_hh_l{l}"] = (3 * d_h, d_h)
        s[f"decoder_rnn.bias_ih_l{l}"] = (3 * d_h,)
        s[f"decoder_rnn.bias_hh_l{l}"] = (3 * d_h,)
    s["decoder_lat.weight"] = (d_h, d_z)
    s["decoder_lat.bias"] = (d_h,)
    s["decoder_fc.weight"] = (V, d_h)
    s["decoder_fc.bias"] = (V,)
    return s


def make_moses_params(seed, dtype=np.float32, **cfg):
    rng = np.random.Generator(np.random.PCG64(seed))
    shapes = moses_shapes(**cfg)
    out = {}
    for k, shp in shapes.items():
        if k == "x_emb.weight":
            # trainable table initialised to the one-hot vectors (mosesvae.py:49-50) plus a small perturbation so
            # that the parity tests see a generic matrix
            out[k] = (np.eye(shp[0]) + 0.05 * rng.standard_normal(shp)).astype(dtype)
            continue
        fan = shp[-1] if len(shp) > 1 else shapes[k.replace("bias", "weight")][-1]
        if "rnn" in k:
            fan = shapes[k.split(".")[0] + ".weight_hh_l0"][1]
        b = 1.0 / np.sqrt(fan)
        out[k] = rng.uniform(-b, b, size=shp).astype(dtype)
    return out


def make_moses_batch(seed, batch, V=34, d_z=160, max_len=100, dtype=np.float32):
    """Length-sorted (desc) list of id arrays with bos/eos: chars are ids 0..V-5,
    specials bos=V-4, eos=V-3, pad=V-2, unk=V-1 (vocab.py:24)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.clip(np.rint(rng.normal(44.0, 9.0, size=batch)), 10, max_len - 2).astype(np.int64)
    lens = np.sort(lens)[::-1]
    bos, eos, pad = V - 4, V - 3, V - 2
    seqs = [np.concatenate([[bos], rng.integers(0, V - 4, size=l), [eos]]).astype(np.int64) for l in lens]
    eps = rng.standard_normal((batch, d_z)).astype(dtype)
    return seqs, eps, pad


def pad_batch(seqs, pad):
    L = np.array([len(s) for s in seqs])
    x = np.full((len(seqs), L.max()), pad, dtype=np.int64)
    for i, s in enumerate(seqs):
        x[i, :len(s)] = s
    return x, L


def _gru_layer_fwd(xin, L, w_ih, w_hh, b_ih, b_hh, h0):
    """xin (B,T,I); packed semantics with lengths L.  Returns outputs (B,T,H) (zero at padded), final h, cache."""
    B, T, _ = xin.shape
    H = w_hh.shape[1]
    gi_all = xin @ w_ih.T + b_ih
    h = h0.copy()
    out = np.zeros((B, T, H), dtype=xin.dtype)
    c = dict(r=[], z=[], n=[], ghn=[], hprev=[], m=[])
    for t in range(T):
        m = (t < L)[:, None]
        gh = h @ w_hh.T + b_hh
        gi = gi_all[:, t]
        r = sigmoid(gi[:, :H] + gh[:, :H])
        z = sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        hn = (1 - z) * n + z * h
        c["r"].append(r); c["z"].append(z); c["n"].append(n); c["ghn"].append(gh[:, 2 * H:]); c["hprev"].append(h); c["m"].append(m)
        h = np.where(m, hn, h)
        out[:, t] = np.where(m, hn, 0.0)
    return out, h, c


def _gru_layer_bwd(dout, dh_final, xin, w_ih, w_hh, c):
    B, T, _ = xin.shape
    H = w_hh.shape[1]
    dgi = np.zeros((B, T, 3 * H), dtype=xin.dtype)
    dgh = np.zeros((B, T, 3 * H), dtype=xin.dtype)
    dh = dh_final.copy()
    for t in reversed(range(T)):
        m = c["m"][t]
        r, z, n, ghn, hp = c["r"][t], c["z"][t], c["n"][t], c["ghn"][t], c["hprev"][t]
        dhn = np.where(m, dh + dout[:, t], 0.0)          # gradient reaching the freshly computed state
        dan = dhn * (1 - z) * (1 - n * n)
        daz = dhn * (hp - n) * z * (1 - z)
        dar = dan * ghn * r * (1 - r)
        dgi[:, t] = np.concatenate([dar, daz, dan], 1)
        dgh[:, t] = np.concatenate([dar, daz, dan * r], 1)
        dh = np.where(m, dhn * z + dgh[:, t] @ w_hh, dh)
    hp_all = np.stack(c["hprev"], 1).reshape(B * T, H)
    g = dict(w_hh=dgh.reshape(B * T, -1).T @ hp_all, b_hh=dgh.sum((0, 1)), b_ih=dgi.sum((0, 1)),
             w_ih=dgi.reshape(B * T, -1).T @ xin.reshape(B * T, -1))
    dx = dgi @ w_ih
    return dx, dh, g


def u01_hash(seed, b, i):
    """The counter-based uniform of molecular-vae_b200/csrc/moses.cu (u01_hash), vectorised over numpy uint64 arrays."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (b.astype(np.uint64) * np.uint64(1000003) + i.astype(np.uint64) + np.uint64(1))) & M
        x ^= x >> np.uint64(30); x = (x * np.uint64(0xBF58476D1CE4E5B9)) & M
        x ^= x >> np.uint64(27); x = (x * np.uint64(0x94D049BB133111EB)) & M
        x ^= x >> np.uint64(31)
    return ((x >> np.uint64(40)).astype(np.float64) + 0.5).astype(np.float32) * np.float32(1.0 / 16777216.0)


def hash64(seed, b, i):
    """The 64-bit value behind u01_hash (same counter construction and mixing)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (b.astype(np.uint64) * np.uint64(1000003) + i.astype(np.uint64) + np.uint64(1))) & M
        x ^= x >> np.uint64(30); x = (x * np.uint64(0xBF58476D1CE4E5B9)) & M
        x ^= x >> np.uint64(27); x = (x * np.uint64(0x94D049BB133111EB)) & M
        x ^= x >> np.uint64(31)
    return x


def dropout_masks(seed, p, B, T, H, d_layers=3, dtype=np.float64):
    """Scaled keep masks (B,T,H) for the outputs of decoder layers 0..L-2: include/mvae_b200.h (mvae_moses_desc.d_dropout).
    One 64-bit hash with counter j // 4 serves four consecutive hidden units, 16 bits each (dropout_kernel in moses.cu)."""
    t, b, j = np.meshgrid(np.arange(T), np.arange(B), np.arange(H), indexing="ij")
    out = []
    for l in range(d_layers - 1):
        h = hash64(seed + l, (t * B + b).astype(np.uint64), (j // 4).astype(np.uint64))
        u16 = (h >> (np.uint64(16) * (j % 4).astype(np.uint64))) & np.uint64(0xFFFF)
        u = (u16.astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 65536.0)
        out.append(((u >= np.float32(p)).astype(dtype) / (1.0 - p)).transpose(1, 0, 2))
    return out


def moses_step(P, seqs, eps, pad, kl_weight=1.0, need_grads=True, d_layers=3, drop_masks=None, dz_ext=None):
    """One fwd(+bwd) step of mosesvae.VAE.forward; the scalar differentiated is kl_weight*kl + recon
    (moses_train_distrib_logp.py:302-306).  dz_ext (B,d_z), optional: gradient wrt z of a loss term computed on z outside
    the VAE (the property head of moses_train_distrib.py:274 / trainbinding.py:216-217), added where autograd would add it."""
    dt = P["decoder_fc.weight"].dtype
    x, L = pad_batch(seqs, pad)
    B, T = x.shape
    E = P["x_emb.weight"]
    V = E.shape[0]
    emb = E[x]                                                  # (B,T,V)
    # ---- encoder (mosesvae.py:150-158)
    q_h = P["encoder_rnn.weight_hh_l0"].shape[1]
    _, h_enc, c_enc = _gru_layer_fwd(emb, L, P["encoder_rnn.weight_ih_l0"], P["encoder_rnn.weight_hh_l0"],
                                     P["encoder_rnn.bias_ih_l0"], P["encoder_rnn.bias_hh_l0"], np.zeros((B, q_h), dt))
    a_mu = h_enc @ P["q_mu.0.weight"].T + P["q_mu.0.bias"]
    r_mu = np.maximum(a_mu, 0)
    mu = r_mu @ P["q_mu.2.weight"].T + P["q_mu.2.bias"]
    a_lv = h_enc @ P["q_logvar.0.weight"].T + P["q_logvar.0.bias"]
    r_lv = np.maximum(a_lv, 0)
    lv = r_lv @ P["q_logvar.2.weight"].T + P["q_logvar.2.bias"]
    std = np.exp(0.5 * lv)
    z = mu + std * eps.astype(dt)
    kl = float((0.5 * (np.exp(lv) + mu ** 2 - 1 - lv).sum(1)).mean(dtype=np.float64))
    # ---- decoder (mosesvae.py:176-197)
    d_h = P["decoder_lat.weight"].shape[0]
    h0 = z @ P["decoder_lat.weight"].T + P["decoder_lat.bias"]
    xin = np.concatenate([emb, np.broadcast_to(z[:, None, :], (B, T, z.shape[1]))], -1)
    caches, inputs = [], []
    cur = xin
    for l in range(d_layers):
        if l >= 1 and drop_masks is not None:
            cur = cur * drop_masks[l - 1]                       # nn.GRU(dropout=p), train mode (mosesvae.py:78)
        inputs.append(cur)
        cur, _, c = _gru_layer_fwd(cur, L, P[f"decoder_rnn.weight_ih_l{l}"], P[f"decoder_rnn.weight_hh_l{l}"],
                                   P[f"decoder_rnn.bias_ih_l{l}"], P[f"decoder_rnn.bias_hh_l{l}"], h0)
        caches.append(c)
    out = cur                                                   # (B,T,d_h), zero at padded positions
    y = out @ P["decoder_fc.weight"].T + P["decoder_fc.bias"]   # (B,T,V)
    tgt = x[:, 1:]
    lg = y[:, :-1]
    mx = lg.max(-1, keepdims=True)
    lse = mx[..., 0] + np.log(np.exp(lg - mx).sum(-1))
    valid = tgt != pad
    M = int(valid.sum())
    nll = lse - np.take_along_axis(lg, tgt[..., None], -1)[..., 0]
    recon = float((nll * valid).sum(dtype=np.float64) / M)
    res = dict(kl=kl, recon=recon, loss=kl_weight * kl + recon, z=z, mu=mu, logvar=lv, y=y, x=x, lengths=L, M=M)
    if not need_grads:
        return res
    G = {k: np.zeros_like(v) for k, v in P.items()}
    # ---- CE backward
    sm = np.exp(lg - lse[..., None])
    dlg = sm.copy()
    np.add.at(dlg, (np.arange(B)[:, None], np.arange(T - 1)[None, :], tgt), -1.0)
    dlg = dlg * valid[..., None] / M
    dy = np.zeros_like(y)
    dy[:, :-1] = dlg
    G["decoder_fc.weight"] = dy.reshape(B * T, V).T @ out.reshape(B * T, d_h)
    G["decoder_fc.bias"] = dy.sum((0, 1))
    dout = dy @ P["decoder_fc.weight"]
    dh0 = np.zeros((B, d_h), dt)
    for l in reversed(range(d_layers)):
        dxl, dh0_l, g = _gru_layer_bwd(dout, np.zeros((B, d_h), dt), inputs[l], P[f"decoder_rnn.weight_ih_l{l}"],
                                       P[f"decoder_rnn.weight_hh_l{l}"], caches[l])
        for nm in ("w_ih", "w_hh", "b_ih", "b_hh"):
            key = {"w_ih": "weight_ih", "w_hh": "weight_hh", "b_ih": "bias_ih", "b_hh": "bias_hh"}[nm]
            G[f"decoder_rnn.{key}_l{l}"] = g[nm]
        dh0 += dh0_l
        dout = dxl if (l == 0 or drop_masks is None) else dxl * drop_masks[l - 1]
    demb = dout[..., :V].copy()
    dz = dout[..., V:].sum(1)
    G["decoder_lat.weight"] = dh0.T @ z
    G["decoder_lat.bias"] = dh0.sum(0)
    dz = dz + dh0 @ P["decoder_lat.weight"]
    if dz_ext is not None:
        dz = dz + dz_ext
    # ---- reparametrisation + KL (weights: kl_weight on kl, 1 on recon)
    dmu = dz + kl_weight * mu / B
    dlv = dz * eps.astype(dt) * std * 0.5 + kl_weight * 0.5 * (np.exp(lv) - 1) / B
    G["q_mu.2.weight"] = dmu.T @ r_mu; G["q_mu.2.bias"] = dmu.sum(0)
    da = (dmu @ P["q_mu.2.weight"]) * (a_mu > 0)
    G["q_mu.0.weight"] = da.T @ h_enc; G["q_mu.0.bias"] = da.sum(0)
    dh_enc = da @ P["q_mu.0.weight"]
    G["q_logvar.2.weight"] = dlv.T @ r_lv; G["q_logvar.2.bias"] = dlv.sum(0)
    da = (dlv @ P["q_logvar.2.weight"]) * (a_lv > 0)
    G["q_logvar.0.weight"] = da.T @ h_enc; G["q_logvar.0.bias"] = da.sum(0)
    dh_enc = dh_enc + da @ P["q_logvar.0.weight"]
    # ---- encoder GRU backward: only the final state receives gradient
    dxe, _, g = _gru_layer_bwd(np.zeros((B, T, q_h), dt), dh_enc, emb, P["encoder_rnn.weight_ih_l0"],
                               P["encoder_rnn.weight_hh_l0"], c_enc)
    G["encoder_rnn.weight_ih_l0"], G["encoder_rnn.weight_hh_l0"] = g["w_ih"], g["w_hh"]
    G["encoder_rnn.bias_ih_l0"], G["encoder_rnn.bias_hh_l0"] = g["b_ih"], g["b_hh"]
    demb = demb + dxe
    dE = np.zeros_like(E)
    np.add.at(dE, x.reshape(-1), demb.reshape(B * T, V))
    dE[pad] = 0.0                                               # nn.Embedding(padding_idx=pad): no gradient on the pad row
    G["x_emb.weight"] = dE
    res["grads"] = G
    return res


def moses_sample_greedy(P, z, bos, eos, pad, max_len=100, d_layers=3, return_margins=False):
    """Greedy restatement of VAE.sample (mosesvae.py:214-262; the shipped uint8 eos_mask / undefined d_z are fixed as
    SURVEY.md 8c prescribes: bool mask, argmax instead of torch.multinomial, ties -> lowest id).  All rows run all
    max_len-1 steps; tokens after EOS are not written; returns ids (B,max_len) filled with pad, lengths end_pads.
    return_margins: the third value is the (B, max_len) top-2 logit margin of every decode step (column 0 = inf), so a
    bit-exactness test can tell a wrong token from a near-tie the oracle's own float64 could have flipped."""
    B = z.shape[0]
    E = P["x_emb.weight"]
    V = E.shape[0]
    d_h = P["decoder_lat.weight"].shape[0]
    h0 = z @ P["decoder_lat.weight"].T + P["decoder_lat.bias"]
    h = [h0.copy() for _ in range(d_layers)]
    w = np.full(B, bos, dtype=np.int64)
    x = np.full((B, max_len), pad, dtype=np.int64)
    x[:, 0] = bos
    end = np.full(B, max_len, dtype=np.int64)
    done = np.zeros(B, dtype=bool)
    margins = np.full((B, max_len), np.inf)
    for i in range(1, max_len):
        inp = np.concatenate([E[w], z], 1)
        for l in range(d_layers):
            H = d_h
            gi = inp @ P[f"decoder_rnn.weight_ih_l{l}"].T + P[f"decoder_rnn.bias_ih_l{l}"]
            gh = h[l] @ P[f"decoder_rnn.weight_hh_l{l}"].T + P[f"decoder_rnn.bias_hh_l{l}"]
            r = sigmoid(gi[:, :H] + gh[:, :H])
            zz = sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            h[l] = (1 - zz) * n + zz * h[l]
            inp = h[l]
        y = inp @ P["decoder_fc.weight"].T + P["decoder_fc.bias"]
        w = y.argmax(-1)
        srt = np.sort(y, -1)
        margins[:, i] = srt[:, -1] - srt[:, -2]
        x[~done, i] = w[~done]
        new_eos = ~done & (w == eos)
        end[new_eos] = i + 1
        done |= new_eos
    return x, end, (margins if return_margins else y)


# ---------------------------------------------------------------------------------------------------------
# mosesfile.py variant (mosesfile.py:6-157): bidirectional encoder GRU (hard-coded, :21-28), single-Linear mu / logvar
# heads on cat(h_fwd, h_bwd) (:31-32, :115-118), d_z = 128 (config.py:34-36), forward returns (kl, recon) (:100).
# Pinned by tests/golden/make_golden_moses.py (mosesfile.VAE with config --q_bidir) -> tests/golden/mosesfile_*.npz.
# ---------------------------------------------------------------------------------------------------------
def mosesfile_shapes(V=34, d_z=128, q_h=256, d_h=512, d_layers=3):
    s = {"x_emb.weight": (V, V)}
    for sfx in ("", "_reverse"):
        s[f"encoder_rnn.weight_ih_l0{sfx}"] = (3 * q_h, V)
        s[f"encoder_rnn.weight_hh_l0{sfx}"] = (3 * q_h, q_h)
        s[f"encoder_rnn.bias_ih_l0{sfx}"] = (3 * q_h,)
        s[f"encoder_rnn.bias_hh_l0{sfx}"] = (3 * q_h,)
    s["q_mu.weight"] = (d_z, 2 * q_h); s["q_mu.bias"] = (d_z,)
    s["q_logvar.weight"] = (d_z, 2 * q_h); s["q_logvar.bias"] = (d_z,)
    for l in range(d_layers):
        inp = V + d_z if l == 0 else d_h
        s[f"decoder_rnn.weight_ih_l{l}"] = (3 * d_h, inp)
        s[f"decoder_rnn.weight_hh_l{l}"] = (3 * d_h, d_h)
        s[f"decoder_rnn.bias_ih_l{l}"] = (3 * d_h,)
        s[f"decoder_rnn.bias_hh_l{l}"] = (3 * d_h,)
    s["decoder_lat.weight"] = (d_h, d_z); s["decoder_lat.bias"] = (d_h,)
    s["decoder_fc.weight"] = (V, d_h); s["decoder_fc.bias"] = (V,)
    return s


def make_mosesfile_params(seed, dtype=np.float32, **cfg):
    rng = np.random.Generator(np.random.PCG64(seed))
    shapes = mosesfile_shapes(**cfg)
    out = {}
    for k, shp in shapes.items():
        if k == "x_emb.weight":
            out[k] = (np.eye(shp[0]) + 0.05 * rng.standard_normal(shp)).astype(dtype)
            continue
        fan = shp[-1] if len(shp) > 1 else shapes[k.replace("bias", "weight")][-1]
        if "rnn" in k:
            fan = shapes[k.split(".")[0] + ".weight_hh_l0"][1]
        b = 1.0 / np.sqrt(fan)
        out[k] = rng.uniform(-b, b, size=shp).astype(dtype)
    return out


def _reverse_valid(a, L):
    """a (B,T,...) -> each row's first L_b entries reversed in place, padding untouched."""
    out = a.copy()
    for b, l in enumerate(L):
        out[b, :l] = a[b, :l][::-1]
    return out


def mosesfile_step(P, seqs, eps, pad, kl_weight=1.0, need_grads=True, d_layers=3):
    """One fwd(+bwd) step of mosesfile.VAE.forward; differentiates kl_weight*kl + recon."""
    dt = P["decoder_fc.weight"].dtype
    x, L = pad_batch(seqs, pad)
    B, T = x.shape
    E = P["x_emb.weight"]
    V = E.shape[0]
    emb = E[x]
    q_h = P["encoder_rnn.weight_hh_l0"].shape[1]
    z0 = np.zeros((B, q_h), dt)
    enc = lambda sfx: (P[f"encoder_rnn.weight_ih_l0{sfx}"], P[f"encoder_rnn.weight_hh_l0{sfx}"],
                       P[f"encoder_rnn.bias_ih_l0{sfx}"], P[f"encoder_rnn.bias_hh_l0{sfx}"])
    _, h_f, c_f = _gru_layer_fwd(emb, L, *enc(""), z0)
    emb_r = _reverse_valid(emb, L)                              # the reverse direction walks t = L_b-1 .. 0
    _, h_b, c_b = _gru_layer_fwd(emb_r, L, *enc("_reverse"), z0)
    h = np.concatenate([h_f, h_b], 1)                           # mosesfile.py:115-116
    mu = h @ P["q_mu.weight"].T + P["q_mu.bias"]
    lv = h @ P["q_logvar.weight"].T + P["q_logvar.bias"]
    std = np.exp(0.5 * lv)
    z = mu + std * eps.astype(dt)
    kl = float((0.5 * (np.exp(lv) + mu ** 2 - 1 - lv).sum(1)).mean(dtype=np.float64))
    d_h = P["decoder_lat.weight"].shape[0]
    h0 = z @ P["decoder_lat.weight"].T + P["decoder_lat.bias"]
    xin = np.concatenate([emb, np.broadcast_to(z[:, None, :], (B, T, z.shape[1]))], -1)
    caches, inputs, cur = [], [], xin
    for l in range(d_layers):
        inputs.append(cur)
        cur, _, c = _gru_layer_fwd(cur, L, P[f"decoder_rnn.weight_ih_l{l}"], P[f"decoder_rnn.weight_hh_l{l}"],
                                   P[f"decoder_rnn.bias_ih_l{l}"], P[f"decoder_rnn.bias_hh_l{l}"], h0)
        caches.append(c)
    out = cur
    y = out @ P["decoder_fc.weight"].T + P["decoder_fc.bias"]
    tgt, lg = x[:, 1:], y[:, :-1]
    mx = lg.max(-1, keepdims=True)
    lse = mx[..., 0] + np.log(np.exp(lg - mx).sum(-1))
    valid = tgt != pad
    M = int(valid.sum())
    nll = lse - np.take_along_axis(lg, tgt[..., None], -1)[..., 0]
    recon = float((nll * valid).sum(dtype=np.float64) / M)
    res = dict(kl=kl, recon=recon, loss=kl_weight * kl + recon, z=z, mu=mu, logvar=lv, y=y, x=x, lengths=L, M=M)
    if not need_grads:
        return res
    G = {k: np.zeros_like(v) for k, v in P.items()}
    sm = np.exp(lg - lse[..., None])
    dlg = sm.copy()
    np.add.at(dlg, (np.arange(B)[:, None], np.arange(T - 1)[None, :], tgt), -1.0)
    dlg = dlg * valid[..., None] / M
    dy = np.zeros_like(y)
    dy[:, :-1] = dlg
    G["decoder_fc.weight"] = dy.reshape(B * T, V).T @ out.reshape(B * T, d_h)
    G["decoder_fc.bias"] = dy.sum((0, 1))
    dout = dy @ P["decoder_fc.weight"]
    dh0 = np.zeros((B, d_h), dt)
    for l in reversed(range(d_layers)):
        dxl, dh0_l, g = _gru_layer_bwd(dout, np.zeros((B, d_h), dt), inputs[l], P[f"decoder_rnn.weight_ih_l{l}"],
                                       P[f"decoder_rnn.weight_hh_l{l}"], caches[l])
        for nm, key in (("w_ih", "weight_ih"), ("w_hh", "weight_hh"), ("b_ih", "bias_ih"), ("b_hh", "bias_hh")):
            G[f"decoder_rnn.{key}_l{l}"] = g[nm]
        dh0 += dh0_l
        dout = dxl
    demb = dout[..., :V].copy()
    dz = dout[..., V:].sum(1)
    G["decoder_lat.weight"] = dh0.T @ z
    G["decoder_lat.bias"] = dh0.sum(0)
    dz = dz + dh0 @ P["decoder_lat.weight"]
    dmu = dz + kl_weight * mu / B
    dlv = dz * eps.astype(dt) * std * 0.5 + kl_weight * 0.5 * (np.exp(lv) - 1) / B
    G["q_mu.weight"] = dmu.T @ h; G["q_mu.bias"] = dmu.sum(0)
    G["q_logvar.weight"] = dlv.T @ h; G["q_logvar.bias"] = dlv.sum(0)
    dh = dmu @ P["q_mu.weight"] + dlv @ P["q_logvar.weight"]
    zeros = np.zeros((B, T, q_h), dt)
    dxe, _, g = _gru_layer_bwd(zeros, dh[:, :q_h], emb, P["encoder_rnn.weight_ih_l0"], P["encoder_rnn.weight_hh_l0"], c_f)
    for nm, key in (("w_ih", "weight_ih"), ("w_hh", "weight_hh"), ("b_ih", "bias_ih"), ("b_hh", "bias_hh")):
        G[f"encoder_rnn.{key}_l0"] = g[nm]
    demb = demb + dxe
    dxr, _, g = _gru_layer_bwd(zeros, dh[:, q_h:], emb_r, P["encoder_rnn.weight_ih_l0_reverse"],
                               P["encoder_rnn.weight_hh_l0_reverse"], c_b)
    for nm, key in (("w_ih", "weight_ih"), ("w_hh", "weight_hh"), ("b_ih", "bias_ih"), ("b_hh", "bias_hh")):
        G[f"encoder_rnn.{key}_l0_reverse"] = g[nm]
    demb = demb + _reverse_valid(dxr, L)
    dE = np.zeros_like(E)
    np.add.at(dE, x.reshape(-1), demb.reshape(B * T, V))
    dE[pad] = 0.0
    G["x_emb.weight"] = dE
    res["grads"] = G
    return res
