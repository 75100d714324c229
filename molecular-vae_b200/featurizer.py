"""Input featurisation.  `OneHotFeaturizer` keeps the reference's host interface (featurizer.py:3-37:
same constructor, method names, return types) and adds `featurize_ids`, the device path: the strings go to the GPU as packed
bytes, `mvae_text_to_ids` turns them into the right-padded u8 id tensor (B, padlength) the model entry points take, and the
(B, T, C) one-hot the reference builds per item on the host (featurizer.py:22-25, data_loader.py:26-31) never exists."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .engine import _p, _stream


class OneHotFeaturizer(object):
    def __init__(self, charset, padlength):
        self.charset = charset
        self.pad_length = padlength
        self._lut = None

    # ---- reference host interface (featurizer.py:8-37), vectorised but value-identical ----
    def featurize(self, smiles):
        return np.array([self.one_hot_encode(smi) for smi in smiles])

    def one_hot_array(self, i):
        return [int(x) for x in [ix == i for ix in range(len(self.charset))]]

    def one_hot_index(self, c):
        return self.charset.index(c)            # ValueError for a character outside the charset, as the reference

    def pad_smi(self, smi):
        return smi.ljust(self.pad_length)

    def one_hot_encode(self, smi):
        idx = np.array([self.one_hot_index(x) for x in self.pad_smi(smi)], dtype=np.int64)
        out = np.zeros((len(idx), len(self.charset)), dtype=np.int64)
        out[np.arange(len(idx)), idx] = 1
        return out

    def one_hot_decode(self, z):
        z1 = []
        for i in range(len(z)):
            s = ''
            for j in range(len(z[i])):
                s += self.charset[int(np.argmax(z[i][j]))]
            z1.append([s.strip()])
        return z1

    def decode_smiles_from_index(self, vec):
        return ''.join(map(lambda x: self.charset[x], vec)).strip()

    # ---- device path ----
    def _tables(self, device):
        if self._lut is None or self._lut.device != device:
            if len(self.charset) > 255:
                raise ValueError("ids are u8: at most 255 charset entries")
            lut = np.full(256, 255, dtype=np.uint8)
            for i, c in enumerate(self.charset):
                b = c.encode("utf-8")
                if len(b) != 1:
                    raise ValueError("the device path needs single-byte charset entries")
                if lut[b[0]] == 255:            # charset.index returns the FIRST match
                    lut[b[0]] = i
            self._lut = torch.from_numpy(lut).to(device)
        return self._lut

    def featurize_ids(self, smiles, device="cuda"):
        """list[str] -> u8 CUDA tensor (N, pad_length) of charset indices, right-padded with the index of ' '
        (= argmax over the last axis of `featurize(smiles)`).  Raises ValueError for a character outside the charset or a
        string longer than pad_length."""
        device = torch.device(device)
        if device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.MvaeError("molecular-vae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        lut = self._tables(device)
        pad_id = self.charset.index(' ')
        enc = [s.encode("utf-8") for s in smiles]
        offs = np.zeros(len(enc) + 1, dtype=np.int32)
        np.cumsum([len(e) for e in enc], out=offs[1:])
        blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
        text = torch.from_numpy(blob.copy()).pin_memory().to(device, non_blocking=True)
        offsets = torch.from_numpy(offs).pin_memory().to(device, non_blocking=True)
        B, T = len(enc), int(self.pad_length)
        ids = torch.empty(B, T, dtype=torch.uint8, device=device)
        bad = torch.zeros(1, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            check(lib.mvae_text_to_ids(_p(text), _p(offsets), B, T, _p(lut), int(pad_id), _p(ids), _p(bad), _stream()))
        flag = int(bad.item())
        if flag & 1:
            raise ValueError("a character is not in the charset")           # featurizer.py:17 (list.index)
        if flag & 2:
            raise ValueError("a string is longer than the pad length")
        return ids
