"""Deterministic synthetic workloads of the shapes BASELINE.json names.

There is no network, so every benchmark and parity input is generated here from a
seed: a Danish-charset word list, a corpus with non-trivial higher-order n-grams,
an n-gram estimator that writes a well-formed (prefix- and suffix-closed) ARPA
file of order 3-6, reference transcripts, "peaky" CTC logits aligned to the
references (trained-model-like) and "flat" logits (random-init-model-like).

Shapes follow the reference: alphabet and ids from R:src/coral/wav2vec2.py:318-322
+ R:config/model/wav2vec2-small.yaml:10 (V = 46, blank = <pad> = 45, "|" = 36);
master seed 4242 from R:config/asr_finetuning.yaml:13; read-aloud durations from
R:DATASET_README.md:75-76 and R:config/evaluation.yaml:9-10; singleton pruning of
orders >= 2 mirrors ``lmplz --prune 0 1 1 ...`` at R:src/coral/ngram.py:127.
"""

from __future__ import annotations

import os

import numpy as np

MASTER_SEED = 4242
CHARACTERS_TO_KEEP = "abcdefghijklmnopqrstuvwxyzæøå0123456789éü"
# tokenizer vocabulary in id order (lower-cased), as R:src/coral/ngram.py:336-339 builds it
CORAL_LABELS = sorted(set(CHARACTERS_TO_KEEP + "|")) + ["<s>", "</s>", "<unk>", "<pad>"]
VOCAB_SIZE = len(CORAL_LABELS)  # 46
BLANK_ID = CORAL_LABELS.index("<pad>")  # 45
DELIM_ID = CORAL_LABELS.index("|")  # 36
_CHAR_TO_ID = {c: i for i, c in enumerate(CORAL_LABELS) if len(c) == 1}
_CHAR_TO_ID[" "] = DELIM_ID

WORD_LETTERS = "abcdefghijklmnopqrstuvwxyzæøåéü"
# rough Danish letter frequencies so that words share prefixes the way real ones do
_LETTER_W = np.array(
    [6.0, 2.0, 0.6, 5.9, 15.5, 2.4, 4.1, 1.6, 6.0, 0.7, 3.4, 5.2, 3.2, 7.2, 4.6, 1.8, 0.1, 9.0,
     5.8, 6.9, 2.0, 2.3, 0.1, 0.1, 0.7, 0.1, 0.9, 0.9, 1.2, 0.2, 0.1]
)


def frames_for_seconds(seconds: float) -> int:
    """wav2vec2 frame count: conv strides 5*2^6 = 320, receptive field 400."""
    return int((int(seconds * 16000) - 400) // 320 + 1)


def audio_seconds(n_frames) -> float:
    """Inverse of the length formula: (T*320 + 80) / 16000."""
    return (np.asarray(n_frames, dtype=np.float64) * 320 + 80) / 16000.0


def _rng(name: str, seed: int = MASTER_SEED) -> np.random.Generator:
    return np.random.default_rng([seed, *[ord(c) for c in name]])


# ------------------------------------------------------------------ words / corpus
def make_word_list(n_words: int = 50_000, seed: int = MASTER_SEED) -> list[str]:
    rng = _rng("words", seed)
    letters = np.array(list(WORD_LETTERS))
    p = _LETTER_W / _LETTER_W.sum()
    words: list[str] = []
    seen: set[str] = set()
    while len(words) < n_words:
        need = n_words - len(words)
        lens = np.minimum(1 + rng.poisson(4.5, size=need * 2), 20)
        flat = rng.choice(len(letters), size=int(lens.sum()), p=p)
        pos = 0
        for L in lens:
            w = "".join(letters[flat[pos : pos + L]])
            pos += L
            if w not in seen:
                seen.add(w)
                words.append(w)
                if len(words) == n_words:
                    break
    return words


class CorpusModel:
    """Zipf unigram draw blended with a sparse random bigram transition table."""

    def __init__(self, n_words: int, seed: int = MASTER_SEED, n_succ: int = 6, p_follow: float = 0.75):
        rng = _rng("transitions", seed)
        ranks = np.arange(1, n_words + 1, dtype=np.float64)
        z = ranks ** -1.1
        self.zipf_cdf = np.cumsum(z / z.sum())
        self.n_words = n_words
        self.p_follow = p_follow
        self.succ = self._zipf(rng, (n_words, n_succ))
        w = np.arange(1, n_succ + 1, dtype=np.float64) ** -1.0
        self.succ_cdf = np.cumsum(w / w.sum())

    def _zipf(self, rng, size):
        return np.minimum(np.searchsorted(self.zipf_cdf, rng.random(size)), self.n_words - 1)

    def sample(self, n_sent: int, name: str, seed: int = MASTER_SEED, min_len=3, max_len=25):
        """Returns (flat word ids int32, sentence lengths int32)."""
        rng = _rng("corpus:" + name, seed)
        lens = rng.integers(min_len, max_len + 1, size=n_sent).astype(np.int32)
        out = np.full((n_sent, max_len), -1, dtype=np.int32)
        cur = self._zipf(rng, n_sent)
        out[:, 0] = cur
        for pos in range(1, max_len):
            follow = rng.random(n_sent) < self.p_follow
            k = np.minimum(np.searchsorted(self.succ_cdf, rng.random(n_sent)), self.succ.shape[1] - 1)
            nxt = np.where(follow, self.succ[cur, k], self._zipf(rng, n_sent))
            out[:, pos] = nxt
            cur = nxt
        mask = np.arange(max_len)[None, :] < lens[:, None]
        return out[mask].astype(np.int32), lens


def sentences_to_text(flat: np.ndarray, lens: np.ndarray, words: list[str]) -> list[str]:
    out = []
    pos = 0
    for L in lens:
        out.append(" ".join(words[i] for i in flat[pos : pos + L]))
        pos += L
    return out


# ------------------------------------------------------------------ ARPA estimator
_MIX = np.uint64(0x9E3779B97F4A7C15)


def _hash_rows(rows: np.ndarray) -> np.ndarray:
    h = np.zeros(rows.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(rows.shape[1]):
            h = (h ^ (rows[:, k].astype(np.uint64) + np.uint64(1))) * _MIX
            h ^= h >> np.uint64(29)
    return h


def estimate_arpa(flat: np.ndarray, lens: np.ndarray, words: list[str], order: int, path: str,
                  discount: float = 0.75) -> dict:
    """Interpolated absolute discounting; n-grams of order >= 2 seen once are pruned.

    Because count(n-gram) <= count(its prefix) and <= count(its suffix), keeping
    count >= 2 yields a prefix- and suffix-closed model (what KenLM's chain
    lookup assumes). Word ids in the file: corpus ids + 3
    (0 <unk>, 1 <s>, 2 </s>).
    """
    UNK, BOS, EOS = 0, 1, 2
    n_sent = len(lens)
    total = int(lens.sum()) + 2 * n_sent
    seq = np.empty(total, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens + 2)[:-1]])
    seq[:] = -1
    seq[starts] = BOS
    seq[starts + lens + 1] = EOS
    seq[seq == -1] = flat.astype(np.int64) + 3
    vocab = ["<unk>", "<s>", "</s>"] + list(words)
    V = len(vocab)

    # unigrams (add-half; <s> gets the conventional -99)
    cnt1 = np.bincount(seq[seq != BOS], minlength=V).astype(np.float64)
    p1 = (cnt1 + 0.5) / (cnt1.sum() + 0.5 * V)
    present1 = cnt1 > 0
    present1[[UNK, BOS, EOS]] = True
    logp = {1: np.log10(p1)}
    logp[1][BOS] = -99.0

    grams = {1: np.arange(V, dtype=np.int64)[:, None]}
    probs = {1: p1.copy()}
    keys = {}
    backoff = {}
    prev_kept_sorted_keys = None
    for n in range(2, order + 1):
        idx = np.arange(total - n + 1)
        win = np.stack([seq[idx + k] for k in range(n)], axis=1)
        ok = np.ones(len(win), dtype=bool)
        for k in range(1, n):
            ok &= win[:, k] != BOS
        win = win[ok]
        h = _hash_rows(win)
        uh, first, cnt = np.unique(h, return_index=True, return_counts=True)
        g = win[first]
        # context statistics over ALL observed n-grams (before pruning)
        ctx_h = _hash_rows(g[:, :-1])
        uctx, inv = np.unique(ctx_h, return_inverse=True)
        ctx_total = np.bincount(inv, weights=cnt.astype(np.float64))
        kept = cnt >= 2
        ctx_kept_types = np.bincount(inv, weights=kept.astype(np.float64), minlength=len(uctx))
        ctx_pruned_mass = np.bincount(inv, weights=np.where(kept, 0, cnt).astype(np.float64), minlength=len(uctx))
        gamma_ctx = (discount * ctx_kept_types + ctx_pruned_mass) / ctx_total
        gk = g[kept]
        ck = cnt[kept].astype(np.float64)
        invk = inv[kept]
        # lower-order probability of the suffix (n-1)-gram (kept by suffix-closure)
        if n == 2:
            p_low = probs[1][gk[:, 1]]
        else:
            sk = _hash_rows(gk[:, 1:])
            pos = np.searchsorted(keys[n - 1], sk)
            pos = np.minimum(pos, len(keys[n - 1]) - 1)
            assert np.all(keys[n - 1][pos] == sk), "suffix closure violated"
            p_low = probs[n - 1][pos]
        p = (ck - discount) / ctx_total[invk] + gamma_ctx[invk] * p_low
        p = np.minimum(p, 1.0)
        hk = uh[kept]
        order_ix = np.argsort(hk)
        keys[n] = hk[order_ix]
        probs[n] = p[order_ix]
        grams[n] = gk[order_ix]
        logp[n] = np.log10(probs[n])
        # back-off of the (n-1)-gram contexts
        if n == 2:
            b = np.zeros(V)
            # unique contexts are unigrams: recover id through any member
            ctx_id = np.zeros(len(uctx), dtype=np.int64)
            ctx_id[inv] = g[:, 0]
            b[ctx_id] = np.log10(gamma_ctx)
            backoff[1] = b
        else:
            b = np.zeros(len(keys[n - 1]))
            pos = np.searchsorted(keys[n - 1], uctx)
            pos = np.minimum(pos, len(keys[n - 1]) - 1)
            hit = keys[n - 1][pos] == uctx
            b[pos[hit]] = np.log10(gamma_ctx[hit])
            backoff[n - 1] = b

    with open(path, "w", encoding="utf-8") as f:
        f.write("\\data\\\n")
        n1 = int(present1.sum())
        f.write(f"ngram 1={n1}\n")
        for n in range(2, order + 1):
            f.write(f"ngram {n}={len(keys[n])}\n")
        f.write("\n\\1-grams:\n")
        b1 = backoff.get(1, np.zeros(V))
        ids = np.nonzero(present1)[0]
        f.write("".join(f"{logp[1][i]:.6f}\t{vocab[i]}\t{b1[i]:.6f}\n" for i in ids))
        for n in range(2, order + 1):
            f.write(f"\n\\{n}-grams:\n")
            gw = grams[n]
            lp = logp[n]
            if n < order:
                bo = backoff.get(n, np.zeros(len(lp)))
                f.write("".join(
                    f"{lp[r]:.6f}\t{' '.join(vocab[t] for t in gw[r])}\t{bo[r]:.6f}\n"
                    for r in range(len(lp))))
            else:
                f.write("".join(
                    f"{lp[r]:.6f}\t{' '.join(vocab[t] for t in gw[r])}\n" for r in range(len(lp))))
        f.write("\n\\end\\\n")
    return {"order": order, "counts": [int(present1.sum())] + [len(keys[n]) for n in range(2, order + 1)]}


# ------------------------------------------------------------------------- logits
def text_to_ids(text: str) -> np.ndarray:
    return np.fromiter((_CHAR_TO_ID[c] for c in text), dtype=np.int32, count=len(text))


def fit_text_to_frames(text: str, n_frames: int) -> str:
    """Drop trailing words until the minimal CTC alignment fits in ``n_frames``."""
    words = text.split(" ")
    while words:
        t = " ".join(words)
        doubles = sum(1 for a, b in zip(t, t[1:]) if a == b)
        if len(t) + doubles <= n_frames:
            return t
        words.pop()
    return ""


def peaky_logits(text: str, n_frames: int, rng: np.random.Generator, peak: float = 12.0,
                 swap: float = 0.05, confusion: float = 0.35) -> np.ndarray:
    """Trained-model-like logits [n_frames, 46] float32 aligned to ``text``.

    ``12 * onehot + N(0, 1)`` alone leaves exactly one token above the -5
    threshold on every frame, which turns beam search into greedy search; real
    acoustic models are confidently wrong now and then. So on a ``confusion``
    fraction of frames a second token (any symbol, or blank) is raised to within
    U(0, 7) of the peak, which keeps 1-3 tokens per frame and tens of live beams.
    """
    ids = text_to_ids(text)
    L = len(ids)
    target = np.full(n_frames, BLANK_ID, dtype=np.int64)
    if L:
        runs = rng.integers(1, 4, size=L)
        gaps = np.zeros(L + 1, dtype=np.int64)  # blanks before char k; gaps[L] trailing
        gaps[1:L][ids[1:] == ids[:-1]] = 1
        need = int(runs.sum() + gaps.sum())
        while need > n_frames:  # shrink runs until it fits (fit_text_to_frames guarantees runs=1 fits)
            k = int(rng.integers(0, L))
            if runs[k] > 1:
                runs[k] -= 1
                need -= 1
        slack = n_frames - need
        if slack:
            add = rng.multinomial(slack, np.full(L + 1, 1.0 / (L + 1)))
            gaps += add
        pos = 0
        for k in range(L):
            pos += gaps[k]
            target[pos : pos + runs[k]] = ids[k]
            pos += runs[k]
    sw = rng.random(n_frames) < swap
    target[sw] = rng.integers(0, 42, size=int(sw.sum()))
    logits = rng.standard_normal((n_frames, VOCAB_SIZE)).astype(np.float32)
    logits[np.arange(n_frames), target] += np.float32(peak)
    cf = np.nonzero(rng.random(n_frames) < confusion)[0]
    if len(cf):
        alt = rng.integers(0, 43, size=len(cf))
        alt[alt == 42] = BLANK_ID
        logits[cf, alt] += (np.float32(peak) - rng.uniform(0.0, 7.0, size=len(cf))).astype(np.float32)
    return logits


def flat_logits(n_frames: int, rng: np.random.Generator, scale: float = 0.5) -> np.ndarray:
    """Random-init-model-like logits: N(0, 0.5^2); ~43 of 46 tokens pass -5."""
    return (rng.standard_normal((n_frames, VOCAB_SIZE)) * scale).astype(np.float32)


def read_aloud_durations(n: int, rng: np.random.Generator) -> np.ndarray:
    """Log-normal with mean 5.87 s clipped to [0.5, 10] s."""
    sigma = 0.45
    mu = np.log(5.87) - sigma * sigma / 2
    return np.clip(rng.lognormal(mu, sigma, size=n), 0.5, 10.0)


def conversation_durations(n: int, rng: np.random.Generator) -> np.ndarray:
    return rng.uniform(5.0, 30.0, size=n)


def corrupt_text(text: str, rng: np.random.Generator, rate: float = 0.07) -> str:
    """i.i.d. char substitution / deletion / insertion at ``rate`` total."""
    alphabet = WORD_LETTERS + " "
    out = []
    r = rng.random(len(text) + 1)
    kind = rng.integers(0, 3, size=len(text) + 1)
    pick = rng.integers(0, len(alphabet), size=len(text) + 1)
    for i, ch in enumerate(text):
        if r[i] < rate:
            if kind[i] == 0:
                out.append(alphabet[pick[i]])
            elif kind[i] == 1:
                continue
            else:
                out.append(alphabet[pick[i]])
                out.append(ch)
        else:
            out.append(ch)
    return "".join(out)


# ----------------------------------------------------------------------- workloads
class Workload:
    """Logits (padded [B, T_max, 46] float32 + lengths), references and the LM path."""

    def __init__(self, logits, lengths, references, arpa_path, labels, name):
        self.logits = logits
        self.lengths = lengths
        self.references = references
        self.arpa_path = arpa_path
        self.labels = labels
        self.name = name

    @property
    def audio_seconds(self) -> float:
        return float(audio_seconds(self.lengths).sum())


def build_lm(cache_dir: str, order: int = 5, n_words: int = 50_000, n_sent: int = 200_000,
             seed: int = MASTER_SEED):
    """Word list + corpus model + ARPA file (cached on disk by parameters)."""
    os.makedirs(cache_dir, exist_ok=True)
    words = make_word_list(n_words, seed)
    model = CorpusModel(n_words, seed)
    path = os.path.join(cache_dir, f"synth_{order}gram_w{n_words}_s{n_sent}_{seed}.arpa")
    if not os.path.exists(path):
        flat, lens = model.sample(n_sent, "train", seed)
        tmp = path + f".tmp{os.getpid()}"
        estimate_arpa(flat, lens, words, order, tmp)
        os.replace(tmp, path)
    return words, model, path


def build_workload(cache_dir: str, n_utts: int, *, order: int = 5, kind: str = "peaky",
                   shape: str = "read_aloud", n_words: int = 50_000, n_sent: int = 200_000,
                   seed: int = MASTER_SEED, name: str = "eval") -> Workload:
    words, model, arpa = build_lm(cache_dir, order, n_words, n_sent, seed)
    rng = _rng(f"workload:{name}:{kind}:{shape}", seed)
    durs = read_aloud_durations(n_utts, rng) if shape == "read_aloud" else conversation_durations(n_utts, rng)
    lengths = np.array([frames_for_seconds(d) for d in durs], dtype=np.int32)
    flat, lens = model.sample(n_utts, "refs:" + name, seed, min_len=3, max_len=25 if shape == "read_aloud" else 60)
    texts = sentences_to_text(flat, lens, words)
    T_max = int(lengths.max())
    logits = np.full((n_utts, T_max, VOCAB_SIZE), -100.0, dtype=np.float32)
    refs = []
    for u in range(n_utts):
        T = int(lengths[u])
        # ~2.2 frames per character leaves room for blanks, like real speech
        t = fit_text_to_frames(texts[u], int(T / 2.2))
        if not t:
            t = texts[u].split(" ")[0][: max(1, T // 3)]
        refs.append(t)
        if kind == "peaky":
            logits[u, :T] = peaky_logits(t, T, rng)
        else:
            logits[u, :T] = flat_logits(T, rng)
    return Workload(logits, lengths, refs, arpa, list(CORAL_LABELS), f"{name}:{kind}:{shape}:{order}gram")
