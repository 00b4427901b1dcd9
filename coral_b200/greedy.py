"""Greedy CTC decoding on the GPU: argmax + collapse, then ids -> strings on the host.

Replaces ``np.argmax`` + ``Wav2Vec2CTCTokenizer.batch_decode`` on the greedy path
(R:src/coral/compute_metrics.py:62-70; HF:models/wav2vec2/tokenization_wav2vec2.py:296-357,
:410-459, :464-530) with ``skip_special_tokens=False`` semantics: repeats are grouped,
tokens equal to the pad token are dropped, the word delimiter becomes ``" "``, every other
token (including ``<s>``, ``</s>``, ``<unk>``) is emitted literally, ids outside the
vocabulary become the unk token, and the joined string is ``strip()``-ed.
"""

from __future__ import annotations

import numpy as np

from . import _lib


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("coral_b200 needs a CUDA device: there is no CPU path")
    return torch


class CTCVocabulary:
    """The id -> token table of a ``Wav2Vec2CTCTokenizer`` (or a plain list of tokens)."""

    def __init__(self, tokens: list[str], pad_id: int, word_delimiter: str = "|", unk_token: str = "<unk>",
                 do_lower_case: bool = False, replace_word_delimiter_char: str = " "):
        self.tokens = list(tokens)
        self.pad_id = int(pad_id)
        self.word_delimiter = word_delimiter
        self.unk_token = unk_token
        self.do_lower_case = do_lower_case
        self.replace_char = replace_word_delimiter_char
        out = [self.replace_char if t == word_delimiter else t for t in self.tokens]
        self._strings = out
        self._single = np.array([len(t) == 1 for t in out], dtype=bool)
        self._cp = np.array([ord(t) if len(t) == 1 else 0 for t in out], dtype=np.uint32)

    @classmethod
    def from_tokenizer(cls, tokenizer) -> "CTCVocabulary":
        vocab = tokenizer.get_vocab()
        size = max(vocab.values()) + 1
        tokens = [tokenizer.unk_token] * size
        for tok, idx in vocab.items():
            tokens[idx] = tok
        return cls(tokens, tokenizer.pad_token_id, tokenizer.word_delimiter_token, tokenizer.unk_token,
                   getattr(tokenizer, "do_lower_case", False),
                   getattr(tokenizer, "replace_word_delimiter_char", " "))

    def to_strings(self, tokens: np.ndarray, lens: np.ndarray) -> list[str]:
        """Collapsed id rows ``[N, T]`` with lengths ``[N]`` -> stripped strings."""
        N, T = tokens.shape
        lens = lens.astype(np.int64)
        mask = np.arange(T)[None, :] < lens[:, None]
        flat = tokens[mask].astype(np.int64)
        off = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        in_range = (flat >= 0) & (flat < len(self._strings))
        if in_range.all() and self._single[flat].all():
            text = self._cp[flat].tobytes().decode("utf-32-le")
            o = off.tolist()
            out = [text[a:b].strip() for a, b in zip(o[:-1], o[1:])]
        else:
            S, unk = self._strings, self.unk_token
            out = []
            for i in range(N):
                row = flat[off[i] : off[i + 1]]
                out.append("".join(S[t] if 0 <= t < len(S) else unk for t in row).strip())
        if self.do_lower_case:
            out = [s.lower() for s in out]
        return out


def greedy_decode_device(logits, lengths=None, blank_id: int = 0, pad_fixup: bool = False, want_ids: bool = False):
    """``logits`` CUDA float32 ``[B, T, V]`` -> device tensors (ids or None, tokens [B, T], lens [B])."""
    torch = _torch()
    dev = logits.device
    B, T, V = logits.shape
    tokens = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
    lens = torch.zeros(B, dtype=torch.int32, device=dev)
    ids = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev) if want_ids else None
    _lib.check(_lib.load().coral_ctc_greedy(
        logits.data_ptr(), lengths.data_ptr() if lengths is not None else None, B, T, V, int(blank_id),
        int(bool(pad_fixup)), ids.data_ptr() if ids is not None else None, tokens.data_ptr(), lens.data_ptr(),
        _lib.stream_ptr(dev)))
    return ids, tokens, lens


def collapse_ids_device(ids, lengths=None, blank_id: int = 0, group_tokens: bool = True):
    torch = _torch()
    dev = ids.device
    B, T = ids.shape
    tokens = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
    lens = torch.zeros(B, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().coral_ctc_collapse(
        ids.data_ptr(), lengths.data_ptr() if lengths is not None else None, B, T, int(blank_id),
        int(bool(group_tokens)), tokens.data_ptr(), lens.data_ptr(), _lib.stream_ptr(dev)))
    return tokens, lens


def _to_device(x, dtype):
    torch = _torch()
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=dev, dtype=dtype, non_blocking=True).contiguous()


def greedy_decode(logits, vocab: CTCVocabulary, lengths=None, pad_fixup: bool = False) -> list[str]:
    """Greedy transcripts of a padded batch ``[B, T, V]`` (numpy or torch, host or device)."""
    torch = _torch()
    d_logits = _to_device(logits, torch.float32)
    d_len = _to_device(lengths, torch.int32) if lengths is not None else None
    _, tokens, lens = greedy_decode_device(d_logits, d_len, vocab.pad_id, pad_fixup)
    return _device_rows_to_strings(tokens, lens, vocab)


def _device_rows_to_strings(tokens, lens, vocab: CTCVocabulary) -> list[str]:
    """Collapsed id rows still on the device -> strings: the vocabulary lookup and the compaction
    run on the GPU and one flat code-point buffer comes back (instead of the padded ``[B, T]`` id
    matrix); multi-character tokens that actually occur fall back to the host path."""
    torch = _torch()
    B, T = tokens.shape
    dev = tokens.device
    table = getattr(vocab, "_d_cp", None)
    if table is None or table.device != dev:
        table = torch.from_numpy(vocab._cp.astype(np.int64)).to(dev).to(torch.int32)
        vocab._d_cp = table
    l64 = lens.to(torch.int64)
    mask = torch.arange(T, device=dev)[None, :] < l64[:, None]
    ids = tokens[mask].to(torch.int64)
    if ids.numel() and (int(ids.min().item()) < 0 or int(ids.max().item()) >= table.numel()):
        return vocab.to_strings(tokens.cpu().numpy(), lens.cpu().numpy())
    flat = table[ids]
    if bool((flat == 0).any().item()):
        return vocab.to_strings(tokens.cpu().numpy(), lens.cpu().numpy())
    off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    torch.cumsum(l64, 0, out=off[1:])
    text = flat.cpu().numpy().view(np.uint32).tobytes().decode("utf-32-le")
    o = off.cpu().tolist()
    out = [text[a:b].strip() for a, b in zip(o[:-1], o[1:])]
    if vocab.do_lower_case:
        out = [s.lower() for s in out]
    return out


def decode_ids(ids, vocab: CTCVocabulary, lengths=None, group_tokens: bool = True) -> list[str]:
    """``tokenizer.batch_decode(ids, group_tokens=...)`` for integer id rows ``[B, T]``."""
    torch = _torch()
    d_ids = _to_device(ids, torch.int32)
    d_len = _to_device(lengths, torch.int32) if lengths is not None else None
    tokens, lens = collapse_ids_device(d_ids, d_len, vocab.pad_id, group_tokens)
    return _device_rows_to_strings(tokens, lens, vocab)
