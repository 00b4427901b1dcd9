"""GPU parity on the named BASELINE configs: every beam of 16 utterances per
cell -- transcripts, order, logit and LM scores, word frames -- against the oracle, for

* config 2's stress variant (flat logits, T up to 499, beam 100),
* config 3 (conversation-shaped, T up to 1499, beam 200, token_min_logp in {-3,-5,-7,-10,-20};
  peaky and flat posteriors),
* config 5 (LM order 3/4/5/6 x beam width 16...512 with the 50k-word LM).

Each cell runs BOTH kernels: the word-frame instantiation behind ``decode_beams_batch`` and the
text-only one the benchmark times (``decode_padded``). tests/cells.py builds the inputs and
serves the oracle's beams (committed goldens for the cells that cost minutes of CPython, a fork
pool in a subprocess otherwise)."""

from __future__ import annotations

import numpy as np
import pytest

import cells
from conftest import beams_equal

pytestmark = pytest.mark.gpu

_DECODERS = {}


def _decoder(labels, arpa):
    from coral_b200.decoder import build_ctcdecoder

    if arpa not in _DECODERS:
        _DECODERS[arpa] = build_ctcdecoder(labels, arpa)
    return _DECODERS[arpa]


def _pad(logits):
    T = max(x.shape[0] for x in logits)
    out = np.full((len(logits), T, logits[0].shape[1]), -100.0, dtype=np.float32)
    for i, x in enumerate(logits):
        out[i, : x.shape[0]] = x
    return out, np.array([x.shape[0] for x in logits], dtype=np.int32)


def _check_cell(name):
    labels, arpa, logits, kw = cells.cell_inputs(name)
    ref, src = cells.oracle_beams(name)
    dec = _decoder(labels, arpa)
    swaps = 0
    # (1) the word-frame kernel: everything pyctcdecode returns per beam
    got = dec.decode_beams_batch(None, logits, **kw)
    assert len(got) == len(ref) == cells.N_UTTS
    for r, g in zip(ref, got):
        swaps += beams_equal(r, g)
    # (2) the text-only kernel (the instantiation bench.py times): all beams, texts and scores
    padded, lengths = _pad(logits)
    bw = kw["beam_width"]
    out = dec.decode_padded(padded, lengths, beam_width=bw, beam_prune_logp=kw["beam_prune_logp"],
                            token_min_logp=kw["token_min_logp"], n_best=bw)
    B, nb = out.lens.shape
    texts = dec.tokens_to_text(out.tokens.reshape(B * nb, -1), out.lens.reshape(-1))
    for u in range(B):
        n = int(out.n_beams[u])
        g = [(texts[u * nb + k], float(out.logit_score[u, k]), float(out.lm_score[u, k])) for k in range(n)]
        swaps += beams_equal(ref[u], g)
    return swaps, src, sum(len(r) for r in ref), max(x.shape[0] for x in logits)


@pytest.fixture(scope="module", autouse=True)
def _oracle_ready():
    """All live oracle passes of this module in ONE subprocess, before the first decode."""
    cells.ensure(list(cells.CELLS))


def test_config2_flat_logits_beam100():
    swaps, src, n, T = _check_cell("c2flat")
    assert T == 499
    print(f"c2flat: {n} oracle beams ({src}), T_max {T}, near-tie swaps {swaps}")


@pytest.mark.parametrize("tml", [3, 5, 7, 10, 20])
def test_config3_conversation_beam200_token_threshold(tml):
    swaps, src, n, T = _check_cell(f"c3_tml{tml}")
    assert T > 1400
    print(f"c3_tml{tml}: {n} oracle beams ({src}), T_max {T}, near-tie swaps {swaps}")


@pytest.mark.parametrize("tml", [3, 5])
def test_config3_flat_logits_beam200(tml):
    swaps, src, n, T = _check_cell(f"c3flat_tml{tml}")
    assert T > 1400
    print(f"c3flat_tml{tml}: {n} oracle beams ({src}), T_max {T}, near-tie swaps {swaps}")


@pytest.mark.parametrize("order", [3, 4, 5, 6])
@pytest.mark.parametrize("beam", [16, 32, 64, 128, 256, 512])
def test_config5_lm_order_x_beam_width(order, beam):
    swaps, src, n, T = _check_cell(f"c5_o{order}_b{beam}")
    print(f"c5 order {order} beam {beam}: {n} oracle beams ({src}), near-tie swaps {swaps}")
