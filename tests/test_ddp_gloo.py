"""world_size-2 gloo test (CPU) of the data-parallel host logic: per-rank gradients of equal shards, averaged through
FlatGradBuffer.allreduce_mean, equal the full-batch gradients (DataParallel semantics, train_distributed.py:87-89).
The per-rank gradients come from the numpy oracle, so no GPU is needed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vae_oracle as vo

Z, H, L, B = 8, 16, 2, 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import molecular_vae_b200 as m
    from molecular_vae_b200.ddp import FlatGradBuffer, PhasedAllReduce, phase_buckets, shard_rows
    P = vo.make_params(3, dtype=np.float64, latent=Z, hidden=H, layers=L)
    ids, onehot, eps = vo.make_batch(4, B, latent=Z, dtype=np.float64)
    lo, hi = shard_rows(B, rank, world)
    r = vo.config_b_step(P, onehot[lo:hi], eps[lo:hi], layers=L)
    keys = m.param_order(L)
    params = [torch.nn.Parameter(torch.from_numpy(P[k]).clone()) for k in keys]
    buf = FlatGradBuffer(params)
    for p, k in zip(params, keys):
        p.grad.copy_(torch.from_numpy(r["grads"][k]))
    if os.environ.get("MVAE_TEST_PHASED") == "1":
        # bucketed exchange in the order the phased step finalises the gradients (top GRU layer + head first)
        buckets = phase_buckets(keys, [p.numel() for p in params], L)
        ar = PhasedAllReduce(buf.flat, buckets)
        for ph in range(L):
            ar.after_phase(ph)
        ar.finish()
    else:
        buf.allreduce_mean()
    loss = torch.tensor([r["loss"]], dtype=torch.float64)
    dist.all_reduce(loss)
    if rank == 0:
        np.savez(os.path.join(out_dir, "ddp.npz"), loss=loss.numpy() / world,
                 **{k: p.grad.numpy() for k, p in zip(keys, params)})
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("phased", ["0", "1"])
def test_allreduce_mean_matches_full_batch(tmp_path, phased, monkeypatch):
    monkeypatch.setenv("MVAE_TEST_PHASED", phased)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "ddp.npz"))
    P = vo.make_params(3, dtype=np.float64, latent=Z, hidden=H, layers=L)
    ids, onehot, eps = vo.make_batch(4, B, latent=Z, dtype=np.float64)
    full = vo.config_b_step(P, onehot, eps, layers=L)
    assert abs(float(got["loss"][0]) - full["loss"]) < 1e-10
    for k, g in full["grads"].items():
        np.testing.assert_allclose(got[k], g, rtol=1e-9, atol=1e-12)


def test_phase_buckets_tile_the_flat_buffer():
    import molecular_vae_b200 as m
    from molecular_vae_b200.ddp import phase_buckets
    shapes = vo.config_b_shapes()
    keys = m.param_order(3)
    numels = [int(np.prod(shapes[k])) for k in keys]
    b = phase_buckets(keys, numels, 3)
    assert len(b) == 3 and b[0][1] == sum(numels) and b[-1][0] == 0
    assert b[1][1] == b[0][0] and b[2][1] == b[1][0]                       # contiguous, no overlap, full cover
    off = dict(zip(keys, np.cumsum([0] + numels[:-1])))
    assert b[0][0] == off["gru.weight_ih_l2"] and b[1][0] == off["gru.weight_ih_l1"]


def test_shard_rows_cover_batch():
    from molecular_vae_b200.ddp import shard_rows
    for n, w in [(4096, 8), (250, 4), (7, 3)]:
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


# ---- MOSES VAE under data parallelism: global token-mean CE and global batch-mean KL ----
def _moses_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from molecular_vae_b200.ddp import FlatGradBuffer, moses_rank_weights
    from oracle import moses_oracle as mo
    P = mo.make_moses_params(5, dtype=np.float64)
    seqs, eps, pad = mo.make_moses_batch(6, 7, dtype=np.float64)
    # deliberately unequal shards (4 + 3 sequences, different target counts): rows rank::world keep each shard sorted
    mine, eps_mine = seqs[rank::world], eps[rank::world]
    n_tgt = sum(len(s) - 1 for s in mine)
    ks, rs = moses_rank_weights(len(mine), n_tgt)
    klw = 0.3
    r = mo.moses_step(P, mine, eps_mine, pad, kl_weight=klw * ks / rs)     # grads of (klw ks kl + rs recon) = rs * these
    keys = sorted(r["grads"].keys())
    params = [torch.nn.Parameter(torch.from_numpy(np.asarray(P[k])).clone()) for k in keys]
    buf = FlatGradBuffer(params)
    for p, k in zip(params, keys):
        p.grad.copy_(torch.from_numpy(rs * r["grads"][k]))
    buf.allreduce_mean()
    if rank == 0:
        np.savez(os.path.join(out_dir, "moses_ddp.npz"), **{k: p.grad.numpy() for k, p in zip(keys, params)})
    dist.destroy_process_group()


def test_moses_rank_weights_reproduce_full_batch_gradients(tmp_path):
    from oracle import moses_oracle as mo
    world = 2
    mp.spawn(_moses_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "moses_ddp.npz"))
    P = mo.make_moses_params(5, dtype=np.float64)
    seqs, eps, pad = mo.make_moses_batch(6, 7, dtype=np.float64)
    full = mo.moses_step(P, seqs, eps, pad, kl_weight=0.3)
    for k, g in full["grads"].items():
        np.testing.assert_allclose(got[k], g, rtol=1e-8, atol=1e-11)


def test_moses_readiness_order_tiles_the_parameters():
    """ddp.moses_readiness_order: every C-ABI key (+ the property head's) exactly once, phase cuts increasing, the top
    decoder layer / decoder_fc / head in phase 0 and decoder layer 0, encoder and x_emb in the last phase."""
    import molecular_vae_b200 as m
    from molecular_vae_b200.mosesfile import mosesfile_param_order
    from molecular_vae_b200.mosesvae import BindingModel, moses_param_order
    head = [f"binding_model.binding_model.{k}" for k in BindingModel.KEYS]
    for keys in (moses_param_order(3), mosesfile_param_order(3), moses_param_order(1)):
        L = 3 if any(k.endswith("_l2") for k in keys) else 1
        for hk in ((), head):
            order, cuts = m.ddp.moses_readiness_order(keys, L, hk)
            assert sorted(order) == sorted(list(keys) + list(hk)) and len(cuts) == L and cuts[-1] == len(order)
            assert all(a < b for a, b in zip(cuts, cuts[1:]))
            first = order[:cuts[0]]
            assert "decoder_fc.weight" in first and f"decoder_rnn.weight_hh_l{L - 1}" in first and all(k in first for k in hk)
            last = order[cuts[-2] if L > 1 else 0:]
            assert "x_emb.weight" in last and "decoder_rnn.weight_ih_l0" in last and "encoder_rnn.weight_hh_l0" in last
