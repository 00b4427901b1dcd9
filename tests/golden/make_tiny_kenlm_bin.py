"""Assembles tests/golden/tiny_bigram.bin byte by byte from KenLM's published probing layout.

Independent of tests/kenlm_binary_writer.py on purpose: that writer and the
reader in coral_b200/csrc/lm_host.cc could share a bug; this script shares no code with either.
Every offset below is written out by hand from the layout of kpu/kenlm (lm/binary_format.cc
`Sanity` + `FixedWidthParameters`, lm/vocab.cc `ProbingVocabulary`, lm/search_hashed.hh,
util/probing_hash_table.hh), for a bigram model over {<unk>, <s>, </s>, a, b}:

  offset  size  field
       0    56  "mmap lm http://kheafield.com/code format version 5\n\0" zero-padded to 56
      56     4  float 0.0          60  4  float 1.0          64  4  float -0.5
      68     4  uint32 1           72  4  uint32 0xFFFFFFFF  76  4  padding
      80     8  uint64 1
      88     1  order = 2 (+3 pad) 92  4  float probing_multiplier = 1.5
      96     4  int32 model_type = 0 (probing)
     100     1  has_vocabulary = 1 (+3 pad)                 104  4  uint32 search_version = 1
     108    16  uint64 counts[2] = {5, 4}
     124     4  padding to a multiple of 8 -> header is 128 bytes
     128     8  ProbingVocabularyHeader {uint32 version = 0, uint32 bound = 5}
     136   112  vocabulary table: max(5 + 1, uint64(1.5f * 5)) = 7 buckets x 16 bytes
                {uint64 MurmurHash64A(word, seed 0), uint32 id, 4 pad} at hash % 7, linear probing;
                "<unk>" is not stored
     248    48  unigrams: (5 + 1) x {float prob, float backoff} in id order + the hallucinated slot
     296    96  bigrams (the longest order): max(4 + 1, uint64(1.5f * 4)) = 6 buckets x 16 bytes
                {uint64 key, float prob, 4 pad}; key = id(word) * 8978948897894561157
                ^ (1 + id(context)) * 17894857484156487943  (mod 2^64), slot = key % buckets
     392    19  words: "<unk>\0<s>\0</s>\0a\0b\0"                      -> 411 bytes

The ARPA twin (tests/golden/tiny_bigram.arpa) holds the same numbers; the test loads both and
requires identical scores.
"""

import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
MASK = 0xFFFFFFFFFFFFFFFF
WORDS = ["<unk>", "<s>", "</s>", "a", "b"]                      # ids 0..4
UNI = [(-1.2, 0.0), (-99.0, -0.30103), (-0.8, 0.0), (-0.5, -0.25), (-0.7, -0.4)]   # log10 prob, backoff
BI = {("<s>", "a"): -0.2, ("a", "b"): -0.3, ("b", "a"): -0.45, ("b", "</s>"): -0.1}  # context, word


def murmur(data: bytes) -> int:  # MurmurHash64A, seed 0 (Austin Appleby, public domain)
    m = 0xC6A4A7935BD1E995
    h = (len(data) * m) & MASK
    body, tail = data[: len(data) // 8 * 8], data[len(data) // 8 * 8:]
    for off in range(0, len(body), 8):
        (k,) = struct.unpack_from("<Q", body, off)
        k = k * m & MASK
        k ^= k >> 47
        k = k * m & MASK
        h = (h ^ k) * m & MASK
    if tail:
        h = (h ^ int.from_bytes(tail, "little")) * m & MASK
    h ^= h >> 47
    h = h * m & MASK
    return h ^ (h >> 47)


def place(table: bytearray, n_buckets: int, key: int, packed: bytes):
    slot = key % n_buckets
    while table[16 * slot: 16 * slot + 8] != b"\0" * 8:
        slot = (slot + 1) % n_buckets
    table[16 * slot: 16 * slot + len(packed)] = packed


def main():
    buf = bytearray(128)
    magic = b"mmap lm http://kheafield.com/code format version 5\n\x00"
    buf[0: len(magic)] = magic
    struct.pack_into("<fff", buf, 56, 0.0, 1.0, -0.5)
    struct.pack_into("<II", buf, 68, 1, 0xFFFFFFFF)
    struct.pack_into("<Q", buf, 80, 1)
    buf[88] = 2
    struct.pack_into("<f", buf, 92, 1.5)
    struct.pack_into("<i", buf, 96, 0)
    buf[100] = 1
    struct.pack_into("<I", buf, 104, 1)
    struct.pack_into("<QQ", buf, 108, len(WORDS), len(BI))
    vocab_buckets = max(len(WORDS) + 1, int(1.5 * len(WORDS)))        # 7
    vocab = bytearray(8 + 16 * vocab_buckets)
    struct.pack_into("<II", vocab, 0, 0, len(WORDS))
    table = bytearray(16 * vocab_buckets)
    for wid, w in enumerate(WORDS):
        if wid:
            h = murmur(w.encode())
            place(table, vocab_buckets, h, struct.pack("<QI", h, wid))
    vocab[8:] = table
    uni = b"".join(struct.pack("<ff", p, b) for p, b in UNI) + struct.pack("<ff", 0.0, 0.0)
    BIGRAM_BUCKETS = max(len(BI) + 1, int(1.5 * len(BI)))              # 6
    bi = bytearray(16 * BIGRAM_BUCKETS)
    for (ctx, w), p in BI.items():
        key = (WORDS.index(w) * 8978948897894561157 & MASK) ^ ((1 + WORDS.index(ctx)) * 17894857484156487943 & MASK)
        place(bi, BIGRAM_BUCKETS, key, struct.pack("<Qf", key, p))
    words = b"".join(w.encode() + b"\0" for w in WORDS)
    with open(os.path.join(HERE, "tiny_bigram.bin"), "wb") as f:
        f.write(bytes(buf) + bytes(vocab) + uni + bytes(bi) + words)
    with open(os.path.join(HERE, "tiny_bigram.arpa"), "w") as f:
        f.write("\\data\\\nngram 1=5\nngram 2=4\n\n\\1-grams:\n")
        for w, (p, b) in zip(WORDS, UNI):
            f.write(f"{p}\t{w}\t{b}\n")
        f.write("\n\\2-grams:\n")
        for (ctx, w), p in BI.items():
            f.write(f"{p}\t{ctx} {w}\n")
        f.write("\n\\end\\\n")


if __name__ == "__main__":
    main()
