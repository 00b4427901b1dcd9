"""Host-side logic on CPU: alphabet normalisation, string marshalling, shims, sharding and
the count all-reduce (gloo, world_size 2), group enumeration of get_score_df."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_alphabet_normalisation_matches_oracle_and_coral_labels():
    import synth
    from coral_b200.alphabet import Alphabet
    from oracle.beam import Alphabet as OracleAlphabet

    a = Alphabet.build_alphabet(synth.CORAL_LABELS)
    assert a.labels == OracleAlphabet.build_alphabet(synth.CORAL_LABELS).labels
    assert a.labels[36] == " " and a.labels[45] == "" and a.labels[44] == "⁇" and a.labels[42] == "<s>"
    assert len(a.labels) == 46 and not a.is_bpe
    assert Alphabet.loads(a.dumps()).labels == a.labels
    assert Alphabet.build_alphabet(["a", "b", "_"]).labels == ["a", "b", ""]
    assert Alphabet.build_alphabet(["a", "b"]).labels == ["a", "b", ""]
    with pytest.raises(NotImplementedError):
        Alphabet.build_alphabet(["▁a", "b"])


def test_text_marshalling_round_trip():
    from coral_b200.textio import decode_utf32, encode_utf32

    strs = ["hej med dig", "", "æøå é ü", "a\tb", "🙂 x"]
    cps, off = encode_utf32(strs)
    assert off.tolist() == [0, 11, 11, 18, 21, 24]
    assert decode_utf32(cps[: off[-1]], off) == strs
    cps, off = encode_utf32([])
    assert off.tolist() == [0] and cps.size >= 1


def test_shims_resolve_to_this_package():
    import coral_b200

    coral_b200.install_shims()
    import kenlm
    import pyctcdecode
    from pyctcdecode import BeamSearchDecoderCTC
    from pyctcdecode.alphabet import BLANK_TOKEN_PTN, UNK_TOKEN, UNK_TOKEN_PTN
    from pyctcdecode.constants import DEFAULT_BEAM_WIDTH, DEFAULT_MIN_TOKEN_LOGP, DEFAULT_PRUNE_LOGP
    from pyctcdecode.decoder import build_ctcdecoder

    from coral_b200 import decoder

    assert BeamSearchDecoderCTC is decoder.BeamSearchDecoderCTC and build_ctcdecoder is decoder.build_ctcdecoder
    assert (DEFAULT_BEAM_WIDTH, DEFAULT_PRUNE_LOGP, DEFAULT_MIN_TOKEN_LOGP) == (100, -10.0, -5.0)
    assert BLANK_TOKEN_PTN.match("<pad>") and UNK_TOKEN_PTN.match("[UNK]") and UNK_TOKEN == "⁇"
    assert BeamSearchDecoderCTC._LANGUAGE_MODEL_SERIALIZED_DIRECTORY == "language_model"
    assert BeamSearchDecoderCTC._ALPHABET_SERIALIZED_FILENAME == "alphabet.json"
    assert hasattr(kenlm, "Model") and pyctcdecode.__file__.startswith(coral_b200.SHIMS_DIR)
    from transformers.utils import is_pyctcdecode_available

    assert is_pyctcdecode_available()


def test_decoder_object_surface_without_gpu():
    """Construction, parameter plumbing and argument errors need no GPU."""
    import synth
    from coral_b200.decoder import BeamSearchDecoderCTC, build_ctcdecoder

    dec = build_ctcdecoder(synth.CORAL_LABELS)
    assert dec._language_model is None and dec.model_container[dec._model_key] is None
    dec.reset_params(alpha=0.7)  # no LM: a no-op, like upstream
    with pytest.raises(ValueError):
        dec._check_logits_dimension(np.zeros((3, 4, 46), np.float32))
    with pytest.raises(ValueError):
        dec._check_logits_dimension(np.zeros((10, 40), np.float32))
    assert dec.tokens_to_text(np.array([[17, 14, 19, 36, 13, 0]], np.uint8), np.array([5])) == ["hej d"]
    import copy

    assert copy.deepcopy(dec) is dec
    dec.cleanup()
    assert dec._model_key not in BeamSearchDecoderCTC.model_container


def test_shard_indices_partition_and_balance():
    from coral_b200.sharded import shard_indices

    rng = np.random.default_rng(0)
    lengths = rng.integers(24, 500, size=1001)
    shards = [shard_indices(lengths, r, 8) for r in range(8)]
    allidx = np.sort(np.concatenate(shards))
    assert np.array_equal(allidx, np.arange(1001))
    loads = [lengths[s].sum() for s in shards]
    assert max(loads) - min(loads) <= 500  # within one utterance of each other


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from coral_b200.sharded import rates_from_totals, reduce_counts, shard_indices
    from oracle import edit

    rng = np.random.default_rng(7)
    refs = ["".join(rng.choice(list("abcde "), size=int(rng.integers(3, 30)))).strip() or "a" for _ in range(101)]
    hyps = [r[: max(1, len(r) - int(rng.integers(0, 3)))] + "x" * int(rng.integers(0, 2)) for r in refs]
    lengths = np.array([len(r) for r in refs])
    groups = rng.integers(0, 3, size=101)
    mine = shard_indices(lengths, rank, world)
    cc = np.array([edit.char_counts(refs[i], hyps[i]) for i in mine], dtype=np.int64).reshape(-1, 4)
    wc = np.array([edit.word_counts(refs[i], hyps[i]) for i in mine], dtype=np.int64).reshape(-1, 4)
    totals = reduce_counts(cc, wc)
    cers, wers = rates_from_totals(totals)
    g_tot = reduce_counts(cc, wc, group_ids=groups[mine], n_groups=3)
    g_cer, _ = rates_from_totals(g_tot)
    q.put((rank, cers[0], wers[0], g_cer, edit.cer(hyps, refs), edit.wer(hyps, refs),
           [edit.cer([hyps[i] for i in np.nonzero(groups == g)[0]], [refs[i] for i in np.nonzero(groups == g)[0]])
            for g in range(3)]))
    dist.destroy_process_group()


def test_sharded_count_all_reduce_gloo_world_size_2():
    """Each rank scores its shard; one all_reduce(SUM) of int64 counts; every rank ends with
    the bit-identical CER/WER of the whole set."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cer, wer, g_cer, ref_cer, ref_wer, ref_g in out:
        assert cer == ref_cer and wer == ref_wer
        assert g_cer == ref_g


def test_group_masks_enumeration_matches_reference_rules():
    import pandas as pd

    from coral_b200.evaluate import group_masks, score_records
    from oracle import edit

    rng = np.random.default_rng(3)
    n = 120
    refs = ["".join(rng.choice(list("abcd "), size=int(rng.integers(3, 25)))).strip() or "a" for _ in range(n)]
    hyps = [r.replace("a", "b", 1) if i % 3 == 0 else r for i, r in enumerate(refs)]
    df = pd.DataFrame(dict(age_group=rng.choice(["0-25", "25-50", "50+"], size=n), gender=rng.choice(["f", "m"], size=n),
                           dialect=rng.choice(["x", "y", "z"], size=n), prediction=hyps, text=refs))
    cats = ["age_group", "gender", "dialect"]
    combos = [c for c, _ in group_masks(df, cats)]
    ref_records = edit.get_score_records(df.to_dict("records"), cats)
    assert combos == [tuple(r[c] for c in cats) for r in ref_records]
    cc = np.array([edit.char_counts(r, h) for r, h in zip(refs, hyps)], dtype=np.int64)
    wc = np.array([edit.word_counts(r, h) for r, h in zip(refs, hyps)], dtype=np.int64)
    assert score_records(df, cats, cc, wc) == ref_records
    # single-valued category: every non-None choice "filters nothing" and is skipped (:187-192)
    df2 = df.assign(gender="f")
    assert all(c[1] is None for c, _ in group_masks(df2, cats))


def test_host_pack_rows_packs_ragged_rows_with_threads():
    """coral_host_pack_rows (pure host code): the list of [T_i, V] arrays becomes one ragged buffer."""
    from coral_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(3)
    for n_threads in (1, 4):
        rows = [rng.standard_normal((int(rng.integers(0, 700)), 46)).astype(np.float32) for _ in range(300)]
        lens = np.array([r.shape[0] for r in rows], dtype=np.int64)
        off = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        ptrs = np.array([r.ctypes.data for r in rows], dtype=np.int64)
        nbytes = lens * 46 * 4
        dst_off = off[:-1] * 46 * 4
        dst = np.full((int(off[-1]), 46), np.nan, dtype=np.float32)
        _lib.check(lib.coral_host_pack_rows(ptrs.ctypes.data, nbytes.ctypes.data, dst_off.ctypes.data, len(rows),
                                            dst.ctypes.data, n_threads))
        assert np.array_equal(dst, np.concatenate(rows, axis=0))
    with pytest.raises(ValueError):
        _lib.check(lib.coral_host_pack_rows(None, None, None, 3, None, 1))


def test_py_logits_rows_reads_the_list_through_the_buffer_protocol():
    """coral_py_logits_rows: addresses and frame counts of float32 C-contiguous [T_i, V] arrays in one C
    loop; everything else (other dtypes, views with strides, wrong width or rank, non-buffers) is marked
    for the caller, and BeamSearchDecoderCTC._rows converts exactly those."""
    import torch

    from coral_b200 import _lib
    from coral_b200.decoder import build_ctcdecoder

    lib = _lib.load()
    rng = np.random.default_rng(5)
    V = 46
    big = rng.standard_normal((9, 50, V)).astype(np.float32)
    items = [
        big[0, :17],                                         # view, contiguous rows
        np.zeros((0, V), np.float32),                        # empty utterance
        big[1, :1],                                          # one frame
        big[2, ::2],                                         # strided rows        -> other
        big[3, :20].astype(np.float64),                      # wrong dtype         -> other
        np.asfortranarray(big[4, :20]),                      # column-major        -> other
        big[5, :20, :45],                                    # wrong width         -> other
        big[6, :20].tolist(),                                # not a buffer        -> other
        torch.from_numpy(big[7, :9]).numpy(),                # round trip through torch
        big[8].reshape(-1),                                  # wrong rank          -> other
    ]
    n = len(items)
    ptrs, lens, other = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.uint8)
    assert lib.coral_py_logits_rows(items, V, ptrs.ctypes.data, lens.ctypes.data, other.ctypes.data) == n
    assert other.tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 0, 1]
    for i in np.nonzero(other == 0)[0]:
        assert lens[i] == items[i].shape[0]
        assert lens[i] == 0 or ptrs[i] == items[i].ctypes.data
    assert lib.coral_py_logits_rows([], V, ptrs.ctypes.data, lens.ctypes.data, other.ctypes.data) == 0

    labels = [chr(ord("a") + i) for i in range(26)] + list("0123456789") + list(" åæéøü") + ["<s>", "</s>", "<unk>", "<pad>"]
    dec = build_ctcdecoder(labels)
    good = [it for k, it in enumerate(items) if k not in (6, 9)]   # wrong width / rank raise, as pyctcdecode does
    p2, l2, keep = dec._rows(good)
    for a, p, l in zip(keep, p2, l2):
        assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous
        assert l == a.shape[0] and (l == 0 or p == a.ctypes.data)
    assert np.array_equal(keep[3], big[2, ::2]) and np.array_equal(keep[4], big[3, :20])
    with pytest.raises(ValueError):
        dec._rows([items[6]])
