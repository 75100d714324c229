"""GPU parity tests of the models.py path ("Config A", cfga.cu) through the C ABI, against oracle/cfga_oracle.py (float64)
and the fixtures generated from the reference's own models.py (tests/golden/cfga_*.npz).

Tolerances (BASELINE.json north_star): fp32 check mode 1e-5 (loss) / 1e-5 relative L2 per gradient tensor; bf16 mode 1e-3
(loss) / 1e-2 (gradients)."""
import os

import numpy as np
import pytest
import torch

from oracle import cfga_oracle as ca
from oracle import vae_oracle as vo
from tests.util_gpu import grads_of, load_pkg, rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(m, P, Z, eh, el, dh, dl, precision):
    mod = m.models
    model = mod.MolecularVAE(i=120, o=Z, c=35, precision=precision)
    model.set_submodules(mod.MolEncoder(i=120, o=Z, c=35, h_size=eh, num_lstm=el),
                         mod.MolDecoder(i=Z, o=120, c=35, num_gru=dl, h_size=dh))
    model.load_state_dict({k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in P.items()}, strict=True)
    return model.cuda()


def _case(ps, bs, B, Z, eh, el, dh, dl):
    P = ca.make_cfga_params(ps, dtype=np.float32, eh=eh, el=el, dh=dh, dl=dl, Z=Z)
    ids, onehot, eps = vo.make_batch(bs, B, latent=Z, dtype=np.float32)
    P64 = {k: v.astype(np.float64) for k, v in P.items()}
    return P, P64, ids, onehot, eps


def _fused(model, ids, eps, use_graph=True):
    out = model.elbo_step(torch.from_numpy(ids).cuda(), torch.from_numpy(eps).cuda(), max_len=120, use_graph=use_graph)
    torch.cuda.synchronize()
    model.engine(ids.shape[0]).check_device_error()
    return out.cpu().numpy()


def _compare(model, sc, ref, loss_rtol, grad_rtol, tag):
    assert abs(sc[0] - ref["loss"]) <= loss_rtol * abs(ref["loss"]), (tag, sc, ref["loss"])
    assert abs(sc[1] - ref["bce"]) <= loss_rtol * abs(ref["bce"]), (tag, sc, ref["bce"])
    assert abs(sc[2] - ref["kl"]) <= loss_rtol * abs(ref["kl"]) + 1e-7, (tag, sc, ref["kl"])
    bad = {k: rel_l2(g, ref["grads"][k]) for k, g in grads_of(model).items()}
    bad = {k: e for k, e in bad.items() if not (e <= grad_rtol)}
    assert not bad, (tag, bad)


@pytest.mark.parametrize("B,Z,eh,el,dh,dl", [(3, 24, 72, 2, 64, 2), (5, 16, 60, 1, 128, 1), (130, 12, 72, 3, 64, 2)])
def test_fused_step_fp32_small(B, Z, eh, el, dh, dl):
    m = load_pkg()
    P, P64, ids, onehot, eps = _case(40 + B, 50 + B, B, Z, eh, el, dh, dl)
    ref = ca.cfga_step(P64, ids.astype(np.int64), eps.astype(np.float64), el=el, dl=dl)
    model = _build(m, P, Z, eh, el, dh, dl, "fp32")
    sc = _fused(model, ids, eps)
    _compare(model, sc, ref, 1e-5, 1e-5, "cfga-fp32-small")
    assert int(sc[3]) == int((ref["argmax"] == ids).all(1).sum())


def test_fused_step_fp32_matches_reference_fixture_small():
    """cfga_small_b3.npz was produced by the reference's models.py itself (tests/golden/make_golden_cfga.py)."""
    m = load_pkg()
    g = np.load(os.path.join(GOLD, "cfga_small_b3.npz"))
    ps, bs, B = [int(v) for v in g["meta"]]
    eh, el, dh, dl, Z = [int(v) for v in g["cfg"]]
    P, _, ids, onehot, eps = _case(ps, bs, B, Z, eh, el, dh, dl)
    model = _build(m, P, Z, eh, el, dh, dl, "fp32")
    sc = _fused(model, ids, eps, use_graph=False)
    assert abs(sc[0] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k, gr in grads_of(model).items():
        gn = float(g[f"gnorm/{k}"])
        if f"gfull/{k}" in g:
            assert rel_l2(gr, g[f"gfull/{k}"]) <= 1e-5, k
        else:
            assert abs(np.sqrt((gr.astype(np.float64) ** 2).sum()) - gn) <= 1e-5 * gn, k
            scale = gn / np.sqrt(gr.size)
            assert np.abs(gr.reshape(-1)[g[f"gidx/{k}"]] - g[f"gval/{k}"]).max() <= 2e-4 * scale, k


def test_fused_step_fp32_full_config_matches_reference_fixture():
    """Full models.py shape (LSTM 3x72 encoder, LSTM 4x1024 decoder, 32.3 M parameters), B=2."""
    m = load_pkg()
    g = np.load(os.path.join(GOLD, "cfga_full_b2.npz"))
    ps, bs, B = [int(v) for v in g["meta"]]
    eh, el, dh, dl, Z = [int(v) for v in g["cfg"]]
    P, _, ids, onehot, eps = _case(ps, bs, B, Z, eh, el, dh, dl)
    model = _build(m, P, Z, eh, el, dh, dl, "fp32")
    assert sum(p.numel() for p in model.parameters()) == 32285105
    sc = _fused(model, ids, eps, use_graph=False)
    assert abs(sc[0] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k, gr in grads_of(model).items():
        gn = float(g[f"gnorm/{k}"])
        assert abs(np.sqrt((gr.astype(np.float64) ** 2).sum()) - gn) <= 2e-5 * gn, (k, gn)
        if f"gfull/{k}" in g:
            assert rel_l2(gr, g[f"gfull/{k}"]) <= 2e-5, k
        else:
            scale = gn / np.sqrt(gr.size)
            assert np.abs(gr.reshape(-1)[g[f"gidx/{k}"]] - g[f"gval/{k}"]).max() <= 5e-4 * scale, k


@pytest.mark.parametrize("B,cell_fused", [(64, "0"), (136, "0"), (64, "1")])
def test_fused_step_bf16_full_config(B, cell_fused, monkeypatch):
    """cell_fused = 1: the LSTM cell runs in the epilogue of the step GEMM (umma_gemm.h mvae_umma_cell, lstm)."""
    monkeypatch.setenv("MVAE_LSTM_CELL_FUSED", cell_fused)
    m = load_pkg()
    P, P64, ids, onehot, eps = _case(71, 72 + B, B, 292, 72, 3, 1024, 4)
    ref = ca.cfga_step(P64, ids.astype(np.int64), eps.astype(np.float64))
    model = _build(m, P, 292, 72, 3, 1024, 4, "bf16")
    sc = _fused(model, ids, eps)
    _compare(model, sc, ref, 1e-3, 1e-2, "cfga-bf16-full")
    sc2 = _fused(model, ids, eps)          # CUDA-graph replay is deterministic up to split-K atomics
    assert abs(sc2[0] - sc[0]) <= 1e-6 * abs(sc[0])


def test_dropin_forward_backward_and_decoder_fp32():
    """model(x) -> (probs, mu, logvar); loss_function; loss.backward() (train.py:98-101); model.decoder(z) (train_sample.py:32)."""
    m = load_pkg()
    B, Z, eh, el, dh, dl = 6, 16, 72, 2, 64, 2
    P, P64, ids, onehot, eps = _case(5, 6, B, Z, eh, el, dh, dl)
    ref = ca.cfga_step(P64, ids.astype(np.int64), eps.astype(np.float64), max_len=128, el=el, dl=dl)
    model = _build(m, P, Z, eh, el, dh, dl, "fp32")
    model.eps_override = torch.from_numpy(eps)
    x = torch.from_numpy(ids.astype(np.int64)).cuda()
    probs, mu, logvar = model(x)
    assert probs.shape == (B, 120, 35) and mu.shape == (B, Z)
    assert model.encoder.lmbd.mu is mu and model.encoder.lmbd.log_v is logvar
    np.testing.assert_allclose(probs.detach().cpu().numpy(), ref["probs"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(mu.detach().cpu().numpy(), ref["mu"], rtol=1e-5, atol=1e-6)
    loss = m.models.loss_function(probs, torch.from_numpy(onehot).cuda(), mu, logvar)   # max_len = 128 as in train.py:43
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    bad = {k: rel_l2(g, ref["grads"][k]) for k, g in grads_of(model).items() if rel_l2(g, ref["grads"][k]) > 1e-5}
    assert not bad, bad
    # decoder on fixed latents: the probabilities the oracle computes from the same z, and their argmax
    z = torch.from_numpy(ref["z"].astype(np.float32)).cuda()
    p2 = model.decoder(z).cpu().numpy()
    np.testing.assert_allclose(p2, ref["probs"], rtol=5e-5, atol=1e-7)
    got = model.decode_greedy(z).cpu().numpy()
    top2 = np.sort(ref["probs"], -1)[..., -2:]
    sure = (top2[..., 1] - top2[..., 0]) > 1e-5
    assert sure.mean() > 0.95 and (got[sure] == ref["argmax"][sure]).all()
