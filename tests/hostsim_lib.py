"""ctypes wrapper around the TEST-ONLY host simulation of the beam kernel (tests/hostsim)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "hostsim", "hostsim.cc"), os.path.join(ROOT, "coral_b200", "csrc", "lm_host.cc")]
DEPS = SRC + [os.path.join(ROOT, "coral_b200", "csrc", f) for f in ("beam_core.h", "lm_tables.h", "lm_host.h")]
OUT = os.path.join(HERE, "hostsim", "_build", "libcoral_hostsim.so")

LOG_BASE_CHANGE = float(1.0 / np.log10(np.e))


def build() -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        subprocess.check_call(
            ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DCORAL_HOSTSIM",
             "-w", *SRC, "-o", OUT]
        )
    return OUT


def _utf32(strings):
    cps = np.array([ord(ch) for s in strings for ch in s], dtype=np.uint32)
    off = np.zeros(len(strings) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in strings])
    return cps, off


class HostSim:
    def __init__(self, labels, arpa_path=None, unigrams="from_arpa"):
        """``labels``: pyctcdecode-normalised alphabet ('' = blank, ' ' = space)."""
        self.lib = C.CDLL(build())
        L = self.lib
        L.hs_create.restype = C.c_void_p
        L.hs_last_error.restype = C.c_char_p
        self.labels = list(labels)
        cps, off = _utf32(self.labels)
        off32 = off.astype(np.int32)
        blank = self.labels.index("")
        space = self.labels.index(" ") if " " in self.labels else -1
        if arpa_path is not None and unigrams == "from_arpa":
            from oracle.arpa import load_unigram_set_from_arpa

            unigrams = sorted(load_unigram_set_from_arpa(arpa_path))
        if unigrams is None or arpa_path is None:
            ucps, uoff, nuni = np.zeros(1, np.uint32), np.zeros(1, np.int64), -1
        else:
            unigrams = list(unigrams)
            ucps, uoff = _utf32(unigrams)
            nuni = len(unigrams)
            if len(ucps) == 0:
                ucps = np.zeros(1, np.uint32)
        self.h = L.hs_create(
            arpa_path.encode() if arpa_path else None,
            cps.ctypes.data_as(C.c_void_p), off32.ctypes.data_as(C.c_void_p), len(self.labels), blank, space,
            ucps.ctypes.data_as(C.c_void_p), uoff.ctypes.data_as(C.c_void_p), C.c_int64(nuni),
        )
        if not self.h:
            raise RuntimeError(L.hs_last_error().decode())
        self.h = C.c_void_p(self.h)

    def score_sentence(self, words, bos=True, eos=True):
        cps, off = _utf32(words)
        if len(cps) == 0:
            cps = np.zeros(1, np.uint32)
        out = np.zeros(len(words) + 1, dtype=np.float32)
        oov = np.zeros(len(words) + 1, dtype=np.int32)
        self.lib.hs_score_sentence(self.h, cps.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                   len(words), int(bos), int(eos), out.ctypes.data_as(C.c_void_p),
                                   oov.ctypes.data_as(C.c_void_p))
        n = len(words) + (1 if eos else 0)
        return out[:n], oov[: len(words)]

    def decode_beams(self, logits, beam_width=100, beam_prune_logp=-10.0, token_min_logp=-5.0, alpha=0.5,
                     beta=1.5, unk_score_offset=-10.0, score_boundary=True, input_mode=0, is_prob=None,
                     variant=0, repeat=1, n_best=None, frames=False, prune_history=False):
        logits = np.ascontiguousarray(logits, dtype=np.float32)
        T, V = logits.shape
        if is_prob is None:
            import math
            is_prob = int(T > 0 and math.isclose(float(logits.sum(axis=1).mean()), 1))
        n_best = beam_width if n_best is None else n_best
        Tm = max(T, 1)
        out_n = np.zeros(1, np.int32)
        out_logit = np.zeros(n_best, np.float64)
        out_comb = np.zeros(n_best, np.float64)
        out_tok = np.zeros((n_best, Tm), np.uint8)
        out_len = np.zeros(n_best, np.int32)
        stats = np.zeros(32, np.uint64)
        self.lib.hs_decode.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
            C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        max_words = (Tm + 1) // 2 + 1
        out_frames = np.full((n_best, max_words, 2), -7, np.int32) if frames else None
        out_nwords = np.zeros(n_best, np.int32) if frames else None
        st = self.lib.hs_decode(
            self.h, logits.ctypes.data, T, int(is_prob), beam_width, beam_prune_logp, token_min_logp, alpha, beta,
            unk_score_offset, int(score_boundary), LOG_BASE_CHANGE, input_mode, n_best,
            variant + (100 if prune_history else 0), repeat,
            out_n.ctypes.data, out_logit.ctypes.data, out_comb.ctypes.data, out_tok.ctypes.data,
            out_len.ctypes.data, stats.ctypes.data,
            out_frames.ctypes.data if frames else None, out_nwords.ctypes.data if frames else None, max_words)
        if st != 0:
            raise RuntimeError(f"hostsim status {st}: {self.lib.hs_last_error().decode()}")
        beams = []
        for r in range(min(int(out_n[0]), n_best)):
            text = "".join(self.labels[t] for t in out_tok[r, : out_len[r]])
            if frames:
                fr = [(int(a), int(b)) for a, b in out_frames[r, : out_nwords[r]]]
                beams.append((text, list(zip(text.split(), fr)), float(out_logit[r]), float(out_comb[r])))
            else:
                beams.append((text, float(out_logit[r]), float(out_comb[r])))
        self.last_stats = stats
        return beams

    def __del__(self):
        try:
            self.lib.hs_free(self.h)
        except Exception:
            pass
