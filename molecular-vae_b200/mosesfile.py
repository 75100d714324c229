"""Drop-in for the reference's mosesfile.py `VAE` (mosesfile.py:6-215) running on the B200 kernels.

`VAE(vocab, config)` with the namespace of config.py:4-85 (q_bidir, q_d_h, q_n_layers, d_z, d_d_h, d_n_layers, d_dropout,
d_cell): the encoder GRU is bidirectional exactly as the reference hard-codes it (mosesfile.py:21-28, hidden 256), the
heads are single Linear layers on cat(h_fwd, h_bwd) (:31-32, :115-118), `forward(list[LongTensor]) -> (kl, recon)` (:100),
`sample(...) -> list[str]` (:215), ModuleList groupings `encoder` (without x_emb, :52-56), `decoder`, `vae` -> identical
state_dict keys.  As in the reference the class only works with `--q_bidir` (otherwise q_mu's input width does not match
the hard-coded bidirectional encoder); that case raises here instead of failing inside a matmul."""
import torch
from torch import nn

from .mosesvae import VAE as _MosesVAE, _check_moses_shapes


def mosesfile_param_order(d_layers=3):
    keys = ["x_emb.weight"]
    for sfx in ("", "_reverse"):
        keys += [f"encoder_rnn.weight_ih_l0{sfx}", f"encoder_rnn.weight_hh_l0{sfx}", f"encoder_rnn.bias_ih_l0{sfx}",
                 f"encoder_rnn.bias_hh_l0{sfx}"]
    keys += ["q_mu.weight", "q_mu.bias", "q_logvar.weight", "q_logvar.bias"]
    for l in range(d_layers):
        keys += [f"decoder_rnn.weight_ih_l{l}", f"decoder_rnn.weight_hh_l{l}", f"decoder_rnn.bias_ih_l{l}",
                 f"decoder_rnn.bias_hh_l{l}"]
    return keys + ["decoder_lat.weight", "decoder_lat.bias", "decoder_fc.weight", "decoder_fc.bias"]


class VAE(_MosesVAE):
    def __init__(self, vocab, config, precision="bf16"):
        nn.Module.__init__(self)
        if not config.q_bidir:
            raise ValueError("mosesfile.VAE: the encoder is bidirectional (mosesfile.py:21-28); pass --q_bidir")
        if config.q_n_layers != 1 or config.d_cell != "gru":
            raise ValueError("mosesfile.VAE on B200 supports q_n_layers=1 and d_cell='gru' (the config.py defaults)")
        self.vocabulary = vocab
        for ss in ("bos", "eos", "unk", "pad"):
            setattr(self, ss, getattr(vocab, ss))
        n_vocab, d_emb = len(vocab), vocab.vectors.size(1)
        if config.q_d_h != 256:
            # mosesfile.py:22-28 hard-codes the encoder GRU's hidden size to 256 while q_mu / q_logvar are sized from
            # config.q_d_h * 2 (:30-32): any other value fails inside the reference's first matmul
            raise ValueError("mosesfile.VAE: config.q_d_h must be 256 (the encoder GRU's hard-coded hidden size, mosesfile.py:22-28)")
        _check_moses_shapes(n_vocab, d_emb, self.pad, 256, config.d_d_h)
        self.x_emb = nn.Embedding(n_vocab, d_emb, self.pad)
        self.x_emb.weight.data.copy_(vocab.vectors)
        self.encoder_rnn = nn.GRU(d_emb, 256, num_layers=config.q_n_layers, batch_first=True, dropout=0.0, bidirectional=True)
        q_d_last = config.q_d_h * 2
        self.q_mu = nn.Linear(q_d_last, config.d_z)
        self.q_logvar = nn.Linear(q_d_last, config.d_z)
        self.decoder_rnn = nn.GRU(d_emb + config.d_z, config.d_d_h, num_layers=config.d_n_layers, batch_first=True,
                                  dropout=config.d_dropout if config.d_n_layers > 1 else 0)
        self.decoder_lat = nn.Linear(config.d_z, config.d_d_h)
        self.decoder_fc = nn.Linear(config.d_d_h, n_vocab)
        self.encoder = nn.ModuleList([self.encoder_rnn, self.q_mu, self.q_logvar])
        self.decoder = nn.ModuleList([self.decoder_rnn, self.decoder_lat, self.decoder_fc])
        self.vae = nn.ModuleList([self.x_emb, self.encoder, self.decoder])
        self.precision = precision
        self.cfg = dict(vocab=n_vocab, d_z=config.d_z, q_hidden=256, d_hidden=config.d_d_h, d_layers=config.d_n_layers,
                        mlp_hidden=256, q_bidir=1, q_linear_heads=1)
        self._keys = mosesfile_param_order(config.d_n_layers)
        self._ws = None
        self.eps_override = None
        self.dropout_seed_override = None

    def forward(self, x):
        """mosesfile.py:84-100: returns (kl_loss, recon_loss)."""
        kl, recon, _, _, _, _ = super().forward(x)
        return kl, recon

    def sample(self, n_batch, max_len=100, z=None, temp=1.0, greedy=False, seed=None):
        """mosesfile.py:168-215: returns list[str]."""
        strs, _ = super().sample(n_batch, max_len, z, temp, greedy, seed)
        return strs
