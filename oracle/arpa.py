"""ARPA n-gram model with KenLM query semantics (oracle; test infrastructure only).

Restates ``kenlm.Model`` / ``kenlm.State`` as the reference reaches them through
pyctcdecode (call sites: R:src/coral/ngram.py:341-343 ``build_ctcdecoder(...,
kenlm_model_path=arpa)``; HF:pipelines/__init__.py:918-933). Upstream:
UP:kenlm ``lm/model.cc`` (``GenericModel::FullScore``, ``ScoreExceptBackoff``,
``ResumeScore``), ``lm/read_arpa.cc``, ``python/kenlm.pyx``. Parity unpinned --
see ``oracle/__init__.py``.

Semantics restated:

* probabilities and back-offs are float32 log10 values;
* ``BaseScore(in, word, out)``: vocabulary index of ``word`` (0 = ``<unk>`` for
  OOV); look up the unigram, then extend the match one context word at a time
  (most recent first) and stop at the first n-gram that is absent
  (``ResumeScore`` returns on ``!pointer.Found()``); ``prob`` is the longest
  match's probability; then ``for i in [ngram_length-1, in.length): prob +=
  in.backoff[i]`` in float32, ascending context length;
* ``out`` = the new word followed by the context words whose n-gram matched,
  with those n-grams' back-offs (KenLM also drops contexts that cannot extend --
  that changes ``State.length`` only, never a score, because a dropped context
  has back-off exactly 0 and no longer n-gram starts with it);
* ``BeginSentenceWrite``: context ``[<s>]`` with ``<s>``'s back-off;
  ``NullContextWrite``: empty context;
* ``word in model``  <=>  vocabulary index != 0.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

F32 = np.float32
UNK_MISSING_LOGPROB = F32(-100.0)  # UP:kenlm lm/config.cc unknown_missing_logprob


@dataclass
class State:
    """KenLM ``State``: context word ids, most recent first, with back-offs."""

    words: tuple[int, ...] = ()
    backoffs: tuple[np.float32, ...] = ()

    def __len__(self) -> int:
        return len(self.words)


@dataclass
class ArpaModel:
    """An ARPA file held as dictionaries, queried with KenLM's rules."""

    path: str
    order: int = 0
    vocab: dict[str, int] = field(default_factory=dict)
    words: list[str] = field(default_factory=list)
    uni_prob: list[np.float32] = field(default_factory=list)
    uni_backoff: list[np.float32] = field(default_factory=list)
    # tables[n] maps (w, c1, c2, ...) -- the predicted word followed by its
    # context, most recent first, n ids in all -- to (prob, backoff)
    tables: dict[int, dict[tuple[int, ...], tuple[np.float32, np.float32]]] = field(
        default_factory=dict
    )
    counts: list[int] = field(default_factory=list)

    # ------------------------------------------------------------------ load
    @classmethod
    def load(cls, path: str) -> "ArpaModel":
        m = cls(path=str(path))
        m.vocab = {"<unk>": 0}
        m.words = ["<unk>"]
        m.uni_prob = [UNK_MISSING_LOGPROB]
        m.uni_backoff = [F32(0.0)]
        saw_unk = False
        section = 0
        declared: dict[int, int] = {}
        pending: list[tuple[int, list[str], np.float32, np.float32]] = []
        with open(path, encoding="utf-8") as f:
            for raw in f:
                line = raw.rstrip("\n").rstrip("\r")
                if not line.strip():
                    continue
                if line.startswith("\\"):
                    tag = line.strip()
                    if tag == "\\data\\":
                        section = 0
                    elif tag == "\\end\\":
                        break
                    elif tag.endswith("-grams:"):
                        section = int(tag[1 : tag.index("-")])
                    else:
                        raise ValueError(f"unknown ARPA section {tag!r}")
                    continue
                if section == 0:
                    if line.startswith("ngram "):
                        n, c = line[6:].split("=")
                        declared[int(n)] = int(c)
                    continue
                parts = line.split("\t")
                if len(parts) < 2:
                    raise ValueError(f"malformed ARPA line {line!r}")
                prob = F32(float(parts[0]))
                toks = parts[1].split(" ")
                if len(toks) != section:
                    raise ValueError(f"expected {section} words in {line!r}")
                backoff = F32(float(parts[2])) if len(parts) > 2 else F32(0.0)
                if section == 1:
                    w = toks[0]
                    if w == "<unk>":
                        if not saw_unk:
                            saw_unk = True
                            m.uni_prob[0] = prob
                            m.uni_backoff[0] = backoff
                    elif w in m.vocab:
                        # duplicate unigram (CoRal's ARPA patch can produce a
                        # second "</s>" line, R:src/coral/ngram.py:150-169): the
                        # first occurrence keeps the vocabulary slot.
                        pass
                    else:
                        m.vocab[w] = len(m.words)
                        m.words.append(w)
                        m.uni_prob.append(prob)
                        m.uni_backoff.append(backoff)
                else:
                    pending.append((section, toks, prob, backoff))
        m.order = max(declared) if declared else max([1] + [p[0] for p in pending])
        m.counts = [declared.get(n, 0) for n in range(1, m.order + 1)]
        for n in range(2, m.order + 1):
            m.tables[n] = {}
        for n, toks, prob, backoff in pending:
            ids = [m.vocab.get(t, 0) for t in toks]
            key = tuple(reversed(ids))  # (w, c1, c2, ...)
            if key in m.tables[n]:
                raise ValueError(f"duplicate {n}-gram {' '.join(toks)!r}")
            m.tables[n][key] = (prob, backoff)
        return m

    # ----------------------------------------------------------------- query
    def index(self, word: str) -> int:
        return self.vocab.get(word, 0)

    def __contains__(self, word: str) -> bool:
        return self.index(word) != 0

    def begin_sentence_state(self) -> State:
        if self.order < 2:
            return State((), ())
        s = self.index("<s>")
        return State((s,), (self.uni_backoff[s],))

    def null_context_state(self) -> State:
        return State((), ())

    def base_score(self, state: State, word: str) -> tuple[float, State]:
        """``BaseScore``: log10 p(word | state) as a Python float, and the out state."""
        return self.base_score_id(state, self.index(word))

    def base_score_id(self, state: State, w: int) -> tuple[float, State]:
        prob = self.uni_prob[w]
        out_words = [w]
        out_backoffs = [self.uni_backoff[w]]
        ngram_length = 1
        key: tuple[int, ...] = (w,)
        for i, cw in enumerate(state.words):
            n = i + 2
            if n > self.order:
                break
            key = key + (cw,)
            hit = self.tables[n].get(key)
            if hit is None:
                break
            prob = hit[0]
            ngram_length = n
            if n < self.order:
                out_words.append(cw)
                out_backoffs.append(hit[1])
        for b in state.backoffs[ngram_length - 1 :]:
            prob = F32(prob + b)
        keep = self.order - 1
        return float(prob), State(tuple(out_words[:keep]), tuple(out_backoffs[:keep]))

    def probes(self, state: State, w: int) -> int:
        """Number of hash slots one scoring touches (for SURVEY 8d's byte count)."""
        return min(self.order, len(state.words) + 1)


def load_unigram_set_from_arpa(arpa_path: str) -> set[str]:
    """UP:pyctcdecode ``language_model.load_unigram_set_from_arpa``:
    the second tab field of every 3-field line between ``\\1-grams:`` and
    ``\\2-grams:``."""
    unigrams: set[str] = set()
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
