"""Test infrastructure: write an ARPA model in KenLM's *probing* binary layout.

KenLM (``UP:kenlm lm/binary_format.cc, lm/vocab.cc, lm/search_hashed.hh, util/probing_hash_table.hh``,
kpu/kenlm master as pinned by ``R:uv.lock:1275-1278``) is not on disk and ``build_binary`` cannot be
run here, so this is a restatement FROM THE PUBLISHED FORMAT, used only to give the reader in
``coral_b200/csrc/lm_host.cc`` (``load_kenlm_binary``) something to read in CI. **Parity unpinned**:
no file produced by the real ``build_binary`` was available to compare byte for byte. The reader
does not trust this writer either -- it validates every structural invariant of a file it is given
(see its header comment) and refuses anything that does not add up.

Layout (little endian):

  Sanity (88 B)        magic "mmap lm http://kheafield.com/code format version 5\\n\\0" padded to 56,
                       float 0, 1, -0.5; uint32 1, 0xFFFFFFFF, 0 (pad); uint64 1
  FixedWidthParameters u8 order (+3 pad), float probing_multiplier, int32 model_type (0 = probing),
                       u8 has_vocabulary (+3 pad), uint32 search_version            (20 B)
  counts               uint64[order]; the whole header is padded to a multiple of 8
  vocabulary           uint32 version, uint32 bound; probing table of {uint64 MurmurHash64A(word),
                       uint32 id, 4 pad}, buckets = max(n + 1, uint64(multiplier * float(n))),
                       slot = hash % buckets, linear probing; "<unk>" (id 0) is not stored
  unigrams             {float prob, float backoff}[n_words + 1]
  middle n = 2..N-1    probing tables of {uint64 key, float prob, float backoff}
  longest              probing table of {uint64 key, float prob, 4 pad}
  words                id order, each NUL-terminated, "<unk>" first

n-gram key: start from the predicted word's id, then for each context word, most recent first,
``key = key * 8978948897894561157 ^ (1 + id) * 17894857484156487943`` (mod 2^64). The sign bit of
a stored middle/unigram probability is a flag ("does not extend left"), the value is -|x|.
"""

from __future__ import annotations

import struct

import numpy as np

MAGIC = b"mmap lm http://kheafield.com/code format version 5\n\x00"
M64 = (1 << 64) - 1


def murmur64a(data: bytes, seed: int = 0) -> int:
    m, r = 0xC6A4A7935BD1E995, 47
    h = (seed ^ (len(data) * m)) & M64
    n8 = len(data) // 8
    for i in range(n8):
        k = int.from_bytes(data[8 * i: 8 * i + 8], "little")
        k = (k * m) & M64
        k ^= k >> r
        k = (k * m) & M64
        h ^= k
        h = (h * m) & M64
    tail = data[8 * n8:]
    if tail:
        h ^= int.from_bytes(tail, "little")
        h = (h * m) & M64
    h ^= h >> r
    h = (h * m) & M64
    h ^= h >> r
    return h


def combine(key: int, word_id: int) -> int:
    return ((key * 8978948897894561157) & M64) ^ (((1 + word_id) * 17894857484156487943) & M64)


def buckets_for(entries: int, multiplier: float) -> int:
    return max(entries + 1, int(np.float32(multiplier) * np.float32(entries)))


def _probing_table(items, n_buckets: int, fmt: str) -> bytes:
    """items: iterable of (key, payload tuple); fmt packs (key, *payload) into one 16-byte bucket."""
    out = bytearray(16 * n_buckets)
    used = [False] * n_buckets
    for key, payload in items:
        assert key != 0, "key 0 marks an empty bucket"
        i = key % n_buckets
        while used[i]:
            i = (i + 1) % n_buckets
        used[i] = True
        struct.pack_into(fmt, out, 16 * i, key, *payload)
    return bytes(out)


def write_probing_binary(model, path: str, multiplier: float = 1.5, flag_some_signs: bool = True) -> None:
    """``model``: ``oracle.arpa.ArpaModel``. ``flag_some_signs`` clears the sign bit of every third
    stored probability, as KenLM does for n-grams that do not extend left (the reader must ignore it)."""
    order = model.order
    n_words = len(model.words)
    counts = [n_words] + [len(model.tables.get(n, {})) for n in range(2, order + 1)]
    head = MAGIC.ljust(56, b"\0") + struct.pack("<fffIIIQ", 0.0, 1.0, -0.5, 1, 0xFFFFFFFF, 0, 1)
    assert len(head) == 88
    head += struct.pack("<B3xfiB3xI", order, multiplier, 0, 1, 1)
    head += struct.pack(f"<{order}Q", *counts)
    head = head.ljust((len(head) + 7) & ~7, b"\0")

    vb = buckets_for(n_words, multiplier)
    vocab = struct.pack("<II", 0, n_words)
    vocab += _probing_table(((murmur64a(w.encode("utf-8")), (i,)) for i, w in enumerate(model.words) if i != 0),
                            vb, "<QI4x")

    def stored(p, k):
        p = -abs(float(p))
        return abs(p) if (flag_some_signs and k % 3 == 0 and p != 0.0) else p

    uni = b"".join(struct.pack("<ff", stored(model.uni_prob[i], i), float(model.uni_backoff[i]))
                   for i in range(n_words))
    uni += struct.pack("<ff", 0.0, 0.0)  # the hallucinated <unk> slot
    search = uni
    for n in range(2, order + 1):
        items = []
        for k, (key_ids, (prob, backoff)) in enumerate(model.tables[n].items()):
            key = key_ids[0]
            for w in key_ids[1:]:
                key = combine(key, w)
            if n < order:
                items.append((key, (stored(prob, k), float(backoff))))
            else:
                items.append((key, (-abs(float(prob)),)))
        nb = buckets_for(len(items), multiplier)
        search += _probing_table(items, nb, "<Qff" if n < order else "<Qf4x")
    words = b"".join(w.encode("utf-8") + b"\0" for w in model.words)
    with open(path, "wb") as f:
        f.write(head + vocab + search + words)
