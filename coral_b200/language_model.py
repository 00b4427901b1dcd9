"""GPU-resident n-gram LM behind pyctcdecode's ``LanguageModel`` / kenlm's ``Model`` surface.

Host-side mirror of UP:pyctcdecode 0.5.0 ``language_model.py`` and the part of
UP:kenlm ``python/kenlm.pyx`` the reference reaches:
``kenlm.Model(path)`` with ``.order`` / ``.path`` / ``in``; ``LanguageModel`` with the
settable ``alpha`` / ``beta`` / ``unk_score_offset`` / ``score_boundary`` attributes that
HF writes through ``decoder.model_container[decoder._model_key]``
(HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:160-183), and
``save_to_dir`` / ``load_from_dir`` with pyctcdecode's file layout. Scoring itself runs on
the device inside the beam-search kernel; there is no host scoring path.
"""

from __future__ import annotations

import ctypes as C
import json
import logging
import os
import shutil

import numpy as np

from . import _lib
from .alphabet import (
    DEFAULT_ALPHA,
    DEFAULT_BETA,
    DEFAULT_SCORE_LM_BOUNDARY,
    DEFAULT_UNK_LOGP_OFFSET,
)
from .textio import encode_utf32

logger = logging.getLogger(__name__)


def _is_kenlm_binary(path: str) -> bool:
    try:
        with open(path, "rb") as f:
            return f.read(33) == b"mmap lm http://kheafield.com/code"
    except OSError:
        return False


def _current_device() -> int:
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("coral_b200 needs a CUDA device: there is no CPU path")
    return torch.cuda.current_device()


class KenlmModel:
    """``kenlm.Model``: an ARPA file or a KenLM probing binary (told apart by the magic bytes, as
    KenLM does) read on the host and held as hash tables in HBM."""

    def __init__(self, path: str, device: int | None = None):
        path = os.fspath(path)
        self.path = path
        self.device = _current_device() if device is None else device
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.coral_lm_load(path.encode(), self.device, C.byref(h)))
        self._h = h
        order = C.c_int32()
        counts = (C.c_uint64 * 8)()
        vocab = C.c_uint64()
        dbytes = C.c_uint64()
        _lib.check(lib.coral_lm_info(h, C.byref(order), counts, C.byref(vocab), C.byref(dbytes)))
        self.order = int(order.value)
        self.ngram_counts = [int(counts[i]) for i in range(self.order)]
        self.vocab_size = int(vocab.value)
        self.device_bytes = int(dbytes.value)
        self.is_binary = _is_kenlm_binary(path)
        if self.is_binary:
            # the reader has never seen a file written by the real build_binary:
            # it refuses anything whose structure does not add up, but say so loudly, and cross-check
            # against the ARPA file when the two sit side by side
            logger.warning(
                "coral_b200: %s was read by a KenLM probing-binary reader that is NOT pinned against a real "
                "build_binary file (none was available to test with). Structural checks passed; if the ARPA "
                "file is at hand, put it next to the .bin (same stem) and it is cross-checked at load time.", path)
            arpa = os.path.splitext(path)[0] + ".arpa"
            if os.path.exists(arpa):
                self._cross_check(arpa)

    def _cross_check(self, arpa_path: str, n: int = 256) -> None:
        """Scores of sentences made of the ARPA's own unigrams must be bit-identical between the
        binary-built and the ARPA-built tables (same float32 numbers, same order of additions)."""
        words = sorted(load_unigram_set_from_arpa(arpa_path) - {"<s>", "</s>", "<unk>"})[: 4 * n]
        if not words:
            return
        sents = [words[i: i + 4] for i in range(0, len(words), 4)]
        other = KenlmModel(arpa_path, self.device)
        a, _ = self.score_sentences(sents)
        b, _ = other.score_sentences(sents)
        for s, x, y in zip(sents, a, b):
            if not np.array_equal(x, y):
                raise _lib.CoralError(_lib.EIO, f"KenLM binary {self.path} disagrees with {arpa_path} on {s}: "
                                                f"{x.tolist()} != {y.tolist()}")
        logger.info("coral_b200: %s agrees with %s on %d sentences", self.path, arpa_path, len(sents))

    def __deepcopy__(self, memo):
        return self

    def contains_batch(self, words) -> np.ndarray:
        words = list(words)
        out = np.zeros(max(len(words), 1), dtype=np.int32)
        if words:
            cps, off = encode_utf32(words)
            _lib.check(_lib.load().coral_lm_contains(self._h, cps.ctypes.data, off.ctypes.data, len(words),
                                                     out.ctypes.data))
        return out[: len(words)].astype(bool)

    def __contains__(self, word: str) -> bool:
        return bool(self.contains_batch([word])[0])

    def score_sentences(self, sentences, bos: bool = True, eos: bool = True):
        """Per-word log10 probabilities (float32) computed on the DEVICE tables -- the
        ``full_scores``-style parity hook for A8. Returns (list of arrays, list of oov arrays)."""
        import torch

        sents = [s.split() if isinstance(s, str) else list(s) for s in sentences]
        words = [w for s in sents for w in s]
        cps, woff = encode_utf32(words)
        soff = np.zeros(len(sents) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in sents], out=soff[1:])
        dev = torch.device("cuda", self.device)
        d_cps = torch.from_numpy(cps.view(np.int32)).to(dev)  # uint32 payload
        d_woff = torch.from_numpy(woff).to(dev)
        d_soff = torch.from_numpy(soff).to(dev)
        n_out = len(words) + (len(sents) if eos else 0)
        d_prob = torch.zeros(max(n_out, 1), dtype=torch.float32, device=dev)
        d_oov = torch.zeros(max(len(words), 1), dtype=torch.int32, device=dev)
        _lib.check(_lib.load().coral_lm_score_sentences(
            self._h, d_cps.data_ptr(), d_woff.data_ptr(), d_soff.data_ptr(), len(sents), int(bos), int(eos),
            d_prob.data_ptr(), d_oov.data_ptr(), _lib.stream_ptr(dev)))
        prob = d_prob.cpu().numpy()
        oov = d_oov.cpu().numpy()
        probs, oovs = [], []
        for i, s in enumerate(sents):
            a = int(soff[i]) + (i if eos else 0)
            probs.append(prob[a : a + len(s) + (1 if eos else 0)].copy())
            oovs.append(oov[int(soff[i]) : int(soff[i + 1])].copy())
        return probs, oovs

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().coral_lm_free(self._h)
                self._h = None
        except Exception:
            pass


def load_unigram_set_from_arpa(arpa_path: str) -> set[str]:
    """UP:pyctcdecode ``language_model.load_unigram_set_from_arpa``."""
    unigrams = set()
    with open(arpa_path, encoding="utf-8") as f:
        start_1_gram = False
        for line in f:
            line = line.strip()
            if line == "\\1-grams:":
                start_1_gram = True
            elif line == "\\2-grams:":
                break
            if start_1_gram and len(line) > 0:
                parts = line.split("\t")
                if len(parts) == 3:
                    unigrams.add(parts[1])
    if len(unigrams) == 0:
        raise ValueError("No unigrams found in arpa file. Something is wrong with the file.")
    return unigrams


class LanguageModel:
    """pyctcdecode ``LanguageModel``: the kenlm model + unigram list + fusion weights."""

    _ATTRS_SERIALIZED_FILENAME = "attrs.json"
    _UNIGRAMS_SERIALIZED_FILENAME = "unigrams.txt"

    def __init__(self, kenlm_model: KenlmModel, unigrams=None, alpha: float = DEFAULT_ALPHA,
                 beta: float = DEFAULT_BETA, unk_score_offset: float = DEFAULT_UNK_LOGP_OFFSET,
                 score_boundary: bool = DEFAULT_SCORE_LM_BOUNDARY) -> None:
        self._kenlm_model = kenlm_model
        if unigrams is None:
            logger.warning("No known unigrams provided, decoding results might be a lot worse.")
            self._unigram_list = None
        else:
            self._unigram_list = sorted(set(unigrams))
        self.alpha = alpha
        self.beta = beta
        self.unk_score_offset = unk_score_offset
        self.score_boundary = score_boundary

    @property
    def order(self) -> int:
        return self._kenlm_model.order

    @property
    def kenlm_model(self) -> KenlmModel:
        return self._kenlm_model

    @property
    def unigrams(self):
        return self._unigram_list

    def save_to_dir(self, filepath: str) -> None:
        os.makedirs(filepath, exist_ok=True)
        with open(os.path.join(filepath, self._ATTRS_SERIALIZED_FILENAME), "w") as fi:
            json.dump({"alpha": self.alpha, "beta": self.beta, "unk_score_offset": self.unk_score_offset,
                       "score_boundary": self.score_boundary}, fi)
        with open(os.path.join(filepath, self._UNIGRAMS_SERIALIZED_FILENAME), "w", encoding="utf-8") as fi:
            fi.write("\n".join(sorted(self._unigram_list or [])))
        dst = os.path.join(filepath, os.path.basename(self._kenlm_model.path))
        if os.path.abspath(dst) != os.path.abspath(self._kenlm_model.path):
            shutil.copy(self._kenlm_model.path, dst)

    @classmethod
    def load_from_dir(cls, filepath: str, unigram_encoding: str | None = None) -> "LanguageModel":
        contents = os.listdir(filepath)
        others = [c for c in contents if c not in (cls._ATTRS_SERIALIZED_FILENAME, cls._UNIGRAMS_SERIALIZED_FILENAME)]
        if len(others) != 1:
            raise ValueError(f"Expected exactly one kenlm model file in {filepath}, found {others}")
        attrs_path = os.path.join(filepath, cls._ATTRS_SERIALIZED_FILENAME)
        json_attrs = json.load(open(attrs_path)) if os.path.exists(attrs_path) else {}
        uni_path = os.path.join(filepath, cls._UNIGRAMS_SERIALIZED_FILENAME)
        unigrams = None
        if os.path.exists(uni_path):
            with open(uni_path, encoding=unigram_encoding or "utf-8") as fi:
                unigrams = fi.read().splitlines()
        return cls(KenlmModel(os.path.join(filepath, others[0])), unigrams, **json_attrs)
