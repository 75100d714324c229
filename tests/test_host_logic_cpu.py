"""CPU checks of the host-side mirror of the reference interface: argument validation the kernels rely on, sequence
packing, the restated loss_function against the reference's own (AST-extracted from baseline/_ref when present)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Vocab:
    def __init__(self, n_chars=30, width=None):
        self.chars = [chr(ord("A") + i) for i in range(n_chars)]
        self.bos, self.eos, self.pad, self.unk = n_chars, n_chars + 1, n_chars + 2, n_chars + 3
        V = n_chars + 4
        self.vectors = torch.eye(V) if width is None else torch.zeros(V, width)

    def __len__(self):
        return len(self.chars) + 4


def test_moses_constructor_validates_what_the_kernels_assume():
    import molecular_vae_b200 as m
    m.mosesvae.VAE(_Vocab())                                        # V = 34, vectors (V, V)
    with pytest.raises(ValueError):
        m.mosesvae.VAE(_Vocab(width=20))                            # embedding width != V: kernels index a V x V table
    with pytest.raises(ValueError):
        m.mosesvae.VAE(_Vocab(n_chars=300))                         # beyond the supported vocabulary

    class Cfg:
        q_cell, q_bidir, q_d_h, q_n_layers, q_dropout = "gru", True, 128, 1, 0.5
        d_cell, d_n_layers, d_dropout, d_z, d_d_h, freeze_embeddings = "gru", 3, 0, 128, 512, False
    with pytest.raises(ValueError):
        m.mosesfile.VAE(_Vocab(), Cfg())                            # mosesfile.py:22-28 hard-codes hidden 256
    Cfg.q_d_h = 256
    m.mosesfile.VAE(_Vocab(), Cfg())


def test_pack_equals_pad_sequence_and_flags_bad_ids():
    import molecular_vae_b200 as m
    model = m.mosesvae.VAE(_Vocab())
    rng = np.random.Generator(np.random.PCG64(1))
    lens = sorted(rng.integers(2, 40, size=50).tolist(), reverse=True)
    x = [torch.from_numpy(rng.integers(0, 34, size=l)) for l in lens]
    x_pad, ids, lens_d = model._pack(x)
    want = torch.nn.utils.rnn.pad_sequence(x, batch_first=True, padding_value=model.pad)
    assert torch.equal(x_pad, want) and torch.equal(ids.long(), want)
    assert lens_d.tolist() == lens and lens_d._host_copy.tolist() == lens
    assert not bool(model._bad_ids)
    with pytest.raises(RuntimeError):
        model._pack(x[::-1])
    x[3] = x[3].clone()
    x[3][1] = 34                                                    # one id outside the vocabulary
    _, ids, _ = model._pack(x)
    assert bool(model._bad_ids) and int(ids.max()) <= 33            # flagged, and clamped for the kernels


def test_loss_function_restatement_equals_the_reference_function():
    """models2d.loss_function / models.loss_function (drop-in) against train.py:31-38 extracted from the reference."""
    from baseline import ref_arm
    if not ref_arm.have_ref():
        pytest.skip("baseline/_ref absent (python baseline/make_ref.py where /root/reference exists)")
    import molecular_vae_b200 as m
    torch.manual_seed(3)
    probs = torch.softmax(torch.randn(7, 120, 35, dtype=torch.float64), -1)
    x = torch.nn.functional.one_hot(torch.randint(0, 35, (7, 120)), 35).double()
    mu, lv = torch.randn(7, 292, dtype=torch.float64), torch.randn(7, 292, dtype=torch.float64)
    for ml in (120, 128):
        ref = ref_arm.load_loss_function(ml)(probs, x, mu, lv)
        m.models2d.max_len = ml
        got = m.models2d.loss_function(probs, x, mu, lv)
        assert abs(float(got) - float(ref)) <= 1e-12 * abs(float(ref))
    m.models2d.max_len = 120


def test_checkpoint_loader_refuses_pickled_code_by_default(tmp_path):
    import molecular_vae_b200 as m

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    path = tmp_path / "ckpt.pt"
    torch.save({"model_state_dict": {}, "epoch": Evil()}, path)
    model = torch.nn.Linear(2, 2)
    with pytest.raises(Exception):
        m.checkpoint.load_reference_checkpoint(model, str(path), strict=False)
