"""ids -> strings on the device: `TokenTable` holds the token texts of a vocabulary / charset on the
GPU; `to_strings` assembles all rows with `mvae_ids_to_text` and brings them to the host in one copy."""
import ctypes

import torch

from . import _lib
from ._lib import check, lib
from .engine import _p, _stream


class TokenTable:
    def __init__(self, tokens, device, rem_first_id=-1, rem_last_id=-1, strip=False):
        """tokens: list of str, text of id 0..len-1 ('' = emit nothing)."""
        if len(tokens) > 256:
            raise ValueError("ids are u8: at most 256 tokens")
        enc = [t.encode("utf-8") for t in tokens]
        self.stride = max(1, max(len(e) for e in enc))
        tab = torch.zeros(256, self.stride, dtype=torch.uint8)
        ln = torch.zeros(256, dtype=torch.uint8)
        for i, e in enumerate(enc):
            if len(e) > 255:
                raise ValueError("token text too long")
            tab[i, :len(e)] = torch.tensor(list(e), dtype=torch.uint8) if e else tab[i, :0]
            ln[i] = len(e)
        self.table, self.tok_len = tab.to(device).contiguous(), ln.to(device)
        self.rem_first_id, self.rem_last_id, self.strip = int(rem_first_id), int(rem_last_id), bool(strip)

    @classmethod
    def from_vocab(cls, vocab, device):
        """vocab.py duck-type: text of id i = ids2string([i]) with nothing removed; bos / eos are dropped only as the
        first / last token of a row (vocab.py:62-73 rem_bos / rem_eos)."""
        toks = [vocab.ids2string([i], rem_bos=False, rem_eos=False) for i in range(len(vocab))]
        return cls(toks, device, rem_first_id=vocab.bos, rem_last_id=vocab.eos)

    def to_strings_async(self, ids, lengths=None):
        """Enqueue the text assembly and the device -> pinned-host copies on the current stream and return a handle whose
        .result() waits for them and builds the list[str] -- so a caller that decodes batch after batch (hugesample.py:27-40)
        can turn batch k into Python strings while the GPU decodes batch k+1."""
        if not ids.is_cuda:
            raise _lib.MvaeError("molecular-vae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        ids = ids.to(torch.uint8).contiguous()
        B, L = ids.shape
        cap = B * L * self.stride
        out = torch.empty(cap, dtype=torch.uint8, device=ids.device)
        offs = torch.empty(B + 1, dtype=torch.int32, device=ids.device)
        scratch = torch.empty(B, dtype=torch.int32, device=ids.device)
        if lengths is not None:
            lengths = lengths.to(device=ids.device, dtype=torch.int32).contiguous()
        with torch.cuda.device(ids.device):
            check(lib.mvae_ids_to_text(_p(ids), _p(lengths), B, L, _p(self.table), self.stride, _p(self.tok_len),
                                       self.rem_first_id, self.rem_last_id, int(self.strip), _p(out), cap, _p(offs),
                                       _p(scratch), _stream()))
            out_h = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            offs_h = torch.empty(B + 1, dtype=torch.int32, pin_memory=True)
            out_h.copy_(out, non_blocking=True)
            offs_h.copy_(offs, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        return _PendingStrings(out_h, offs_h, ev, (out, offs, scratch, ids, lengths))

    def to_strings(self, ids, lengths=None):
        """ids: u8 CUDA tensor (B,L); lengths: int32 CUDA tensor (B) or None.  Returns list[str]."""
        if not ids.is_cuda:
            raise _lib.MvaeError("molecular-vae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        ids = ids.to(torch.uint8).contiguous()
        B, L = ids.shape
        cap = B * L * self.stride
        out = torch.empty(cap, dtype=torch.uint8, device=ids.device)
        offs = torch.empty(B + 1, dtype=torch.int32, device=ids.device)
        scratch = torch.empty(B, dtype=torch.int32, device=ids.device)
        if lengths is not None:
            lengths = lengths.to(device=ids.device, dtype=torch.int32).contiguous()
        with torch.cuda.device(ids.device):
            check(lib.mvae_ids_to_text(_p(ids), _p(lengths), B, L, _p(self.table), self.stride, _p(self.tok_len),
                                       self.rem_first_id, self.rem_last_id, int(self.strip), _p(out), cap, _p(offs),
                                       _p(scratch), _stream()))
        offs_h = offs.cpu().tolist()
        data = bytes(out[:offs_h[-1]].cpu().numpy())
        return [data[offs_h[b]:offs_h[b + 1]].decode("utf-8") for b in range(B)]


class _PendingStrings:
    def __init__(self, out_h, offs_h, event, keep):
        self.out_h, self.offs_h, self.event, self._keep = out_h, offs_h, event, keep
        self.d2h_bytes = out_h.numel() + offs_h.numel() * 4

    def result(self):
        self.event.synchronize()
        self._keep = None
        offs = self.offs_h.tolist()
        data = self.out_h.numpy()[:offs[-1]].tobytes()
        return [data[offs[b]:offs[b + 1]].decode("utf-8") for b in range(len(offs) - 1)]
