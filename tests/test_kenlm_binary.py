"""KenLM probing binary reader on the CPU: a model written in the
published layout by tests/kenlm_binary_writer.py must load and score bit for bit like the ARPA
file it came from, decode identically, and damaged or unsupported files must be refused.
PARITY UNPINNED: no output of the real build_binary is available here (see the writer's header)."""

from __future__ import annotations

import os
import struct

import numpy as np
import pytest

from conftest import beams_equal


@pytest.fixture(scope="module")
def small_bin(small_lm, tmp_path_factory):
    from kenlm_binary_writer import write_probing_binary
    from oracle.arpa import ArpaModel

    path = str(tmp_path_factory.mktemp("klm") / "small.bin")
    write_probing_binary(ArpaModel.load(small_lm[2]), path)
    return path


def test_murmur_and_bucket_rule():
    from kenlm_binary_writer import buckets_for, murmur64a

    # MurmurHash64A reference values (seed 0) computed from the published algorithm
    assert murmur64a(b"") == 0
    assert murmur64a(b"<unk>") == 0xEA91E7561F5B392B
    assert murmur64a(b"hello world") == 0xD3BA2368A832AFCE
    assert buckets_for(10, 1.5) == 15 and buckets_for(1, 1.5) == 2 and buckets_for(0, 1.5) == 1


def test_binary_scores_bit_exact_and_decodes_identically(small_bin, small_lm, oracle_decoder, small_workload, rng):
    import synth
    from hostsim_lib import HostSim

    words, model, arpa = small_lm
    hb = HostSim(oracle_decoder._alphabet.labels, small_bin, unigrams=sorted(oracle_decoder._language_model._unigram_set))
    ha = HostSim(oracle_decoder._alphabet.labels, arpa)
    m = oracle_decoder._language_model._kenlm_model
    flat, lens = model.sample(120, "klm-bin")
    for s in synth.sentences_to_text(flat, lens, words):
        ws = s.split(" ")
        if rng.random() < 0.4:
            ws[int(rng.integers(len(ws)))] = "qqzzx"
        pb, ob = hb.score_sentence(ws)
        pa, oa = ha.score_sentence(ws)
        assert np.array_equal(pb.view(np.uint32), pa.view(np.uint32)) and np.array_equal(ob, oa)
        st = m.begin_sentence_state()
        for w, got in zip(ws, pb):
            want, st = m.base_score(st, w)
            assert np.float32(want) == got
    w = small_workload
    for u in range(3):
        lg = w.logits[u, : w.lengths[u]]
        beams_equal(oracle_decoder.decode_beams(lg), hb.decode_beams(lg, frames=True))


def test_damaged_and_unsupported_binaries_are_refused(small_bin, oracle_decoder, tmp_path):
    from hostsim_lib import HostSim

    labels = oracle_decoder._alphabet.labels
    data = bytearray(open(small_bin, "rb").read())

    def refuses(mut, needle):
        p = str(tmp_path / "x.bin")
        open(p, "wb").write(bytes(mut))
        with pytest.raises(RuntimeError) as e:
            HostSim(labels, p, unigrams=None)
        assert needle in str(e.value), str(e.value)

    t = bytearray(data); struct.pack_into("<i", t, 96, 2)          # model type trie
    refuses(t, "trie")
    t = bytearray(data); t[100] = 0                                # no vocabulary
    refuses(t, "vocabulary")
    t = bytearray(data); struct.pack_into("<f", t, 60, 2.0)        # sanity constant
    refuses(t, "sanity")
    t = bytearray(data); t[49] = ord("4")                          # format version
    refuses(t, "version")
    refuses(data[: len(data) - 7], "")                             # truncated word list
    order = data[88]
    header = (108 + 8 * order + 7) & ~7
    t = bytearray(data); t[header + 8 + 3] ^= 0x55                 # a vocabulary hash
    refuses(t, "")
    c0 = struct.unpack_from("<Q", data, 108)[0]
    t = bytearray(data); struct.pack_into("<Q", t, 108 + 8, struct.unpack_from("<Q", data, 116)[0] + 1)  # bigram count
    refuses(t, "")
    assert c0 > 0


def test_python_wrapper_sniffs_the_magic(small_bin):
    """coral_b200.language_model.KenlmModel picks the loader by content, like kenlm.Model."""
    import ctypes

    from coral_b200 import _lib

    lib = _lib.load()
    for name in ("coral_lm_load", "coral_lm_load_kenlm_binary"):
        assert hasattr(lib, name)
    h = ctypes.c_void_p()
    with pytest.raises((_lib.CoralError, OSError, RuntimeError, ValueError)):
        _lib.check(lib.coral_lm_load_kenlm_binary(b"/nonexistent/file.bin", 0, ctypes.byref(h)))
    assert os.path.getsize(small_bin) > 1000


def test_hand_assembled_binary_fixture_reads_like_its_arpa_twin():
    """tests/golden/tiny_bigram.bin was assembled byte by byte from KenLM's published layout by
    tests/golden/make_tiny_kenlm_bin.py, which shares no code with tests/kenlm_binary_writer.py or
    with the reader: the reader must accept it and score exactly like the ARPA
    twin, OOV words and both start states included. (Still not a file from the real build_binary.)"""
    import os

    from hostsim_lib import HostSim

    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    assert os.path.getsize(os.path.join(g, "tiny_bigram.bin")) == 411
    labels = ["", " ", "a", "b"]
    arpa = HostSim(labels, os.path.join(g, "tiny_bigram.arpa"))
    binary = HostSim(labels, os.path.join(g, "tiny_bigram.bin"), unigrams=["a", "b"])
    for s in (["a", "b", "a"], ["b", "b"], ["a", "zz", "b"], ["b"], []):
        for bos in (True, False):
            x, ox = arpa.score_sentence(s, bos=bos)
            y, oy = binary.score_sentence(s, bos=bos)
            assert np.array_equal(x, y) and ox.tolist() == oy.tolist(), (s, bos)
    got, _ = binary.score_sentence(["a", "b", "a"])
    # hand-computed: p(a|<s>) = -0.2, p(b|a) = -0.3, p(a|b) = -0.45, p(</s>|a) = backoff(a) + p(</s>) = -0.25 - 0.8
    assert np.allclose(got, [-0.2, -0.3, -0.45, -1.05], atol=1e-6)
