"""BASELINE.json configs 3, 4 and 5 as bench lines (``python bench.py --config {3,4,5}``).

The driver's default run is config 2 (bench.py). These are the other named workloads, launched
the same way (``--gpus N`` under torchrun, one rank per GPU) and printing ONE JSON line each with
the same keys. They are parity-gated in tests/test_gpu_configs.py (all beams of 16 utterances per
cell) and tests/test_gpu_parity.py (config 4 properties); here a small oracle sample gates the line.

config 3  conversation-shaped utterances (T up to 1499), beam 200, token_min_logp sweep
          {-3,-5,-7,-10,-20}; value = utterances/s at -5 on peaky logits; weak scaling
config 4  dataset_validation scoring: 200 000 (reference, hypothesis) pairs, per-sample CER / WER,
          keep-mask cer < 0.6 (R:config/dataset_validation.yaml:21, R:src/coral/validation.py:136-159),
          aggregate CER / WER through the all-reduce; STRONG scaling: the 200 000 pairs are split
          over the ranks with shard_indices
config 5  LM order {3,4,5,6} x beam width {16...512} on 2 048 read-aloud utterances; value =
          utterances/s at order 5, beam 128; weak scaling
"""

from __future__ import annotations

import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _setup(local_rank, world):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, dev


def _max_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _time_launches(torch, fn, steps, warmup):
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650"


def _oracle_gate(wl, dec, kw, n=8):
    """Transcripts of the n shortest utterances == the oracle's (a fork pool before CUDA is touched is
    not possible here -- CUDA is up -- so this runs sequentially on a few short utterances)."""
    from oracle.beam import build_ctcdecoder as oracle_build

    o = oracle_build(wl.labels, wl.arpa_path)
    idx = np.argsort(wl.lengths, kind="stable")[:n]
    lg = [wl.logits[u, : wl.lengths[u]] for u in idx]
    got = dec.decode_beams_batch(None, lg, n_best=1, **kw)
    ref = [o.decode_beams(x, **kw)[0][0] for x in lg]
    ok = [g[0][0] for g in got] == ref
    if not ok:
        raise SystemExit("PARITY FAILURE against the oracle: " + json.dumps(kw))
    return {"transcripts_identical": True, "sample": int(n), "kwargs": kw}


def _common(metric, unit, world, steps, warmup, ms, scaling):
    return {"metric": metric, "unit": unit, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "data": "synthetic"}


# ------------------------------------------------------------------------------------ config 3
def run_config3(args, rank, world, local_rank, clock_sampler):
    import synth
    from coral_b200 import metrics
    from coral_b200.decoder import build_ctcdecoder

    torch, dist, dev = _setup(local_rank, world)
    n = args.utts if args.utts != 8192 else 1024
    wl = synth.build_workload(os.environ.get("CORAL_B200_CACHE", "/tmp/coral_b200_cache"), n, order=5, kind="peaky",
                              shape="conversation", name=f"conv{rank}")
    dec = build_ctcdecoder(wl.labels, wl.arpa_path)
    gate = _oracle_gate(wl, dec, dict(beam_width=200, token_min_logp=-5.0), n=4) if rank == 0 else None
    d_logits = torch.from_numpy(wl.logits).to(dev)
    d_len = torch.from_numpy(wl.lengths).to(dev)
    d_order = torch.argsort(d_len, descending=True).to(torch.int32)
    h_logits = torch.from_numpy(wl.logits).pin_memory()
    sampler = clock_sampler(local_rank)
    if rank == 0:
        sampler.start()
    t0c = time.time()
    sweep = {}
    for tml in (-3.0, -5.0, -7.0, -10.0, -20.0):
        steps = args.steps if tml > -15 else max(1, args.steps // 3)
        ms = _time_launches(torch, lambda: dec.decode_launch(d_logits, d_len, d_order, beam_width=200, token_min_logp=tml),
                            steps, args.warmup if tml > -15 else 1)
        ms = _max_over_ranks(ms, dev, world)
        sweep[str(tml)] = {"ms": ms, "utt_per_s": world * n / ms * 1e3}
    # end to end at -5: pinned host logits -> strings -> cer/wer
    def e2e():
        hyps = dec.decode_batch(None, h_logits, beam_width=200, token_min_logp=-5.0, lengths=wl.lengths)
        return metrics.cer(hyps, wl.references), metrics.wer(hyps, wl.references)
    for _ in range(2):
        e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cer_v, wer_v = e2e()
    torch.cuda.synchronize()
    e2e_ms = _max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3, dev, world)
    # flat posteriors (the stress variant), 128 utterances
    fl = synth.build_workload(os.environ.get("CORAL_B200_CACHE", "/tmp/coral_b200_cache"), 128, order=5, kind="flat",
                              shape="conversation", name=f"convflat{rank}")
    f_logits = torch.from_numpy(fl.logits).to(dev)
    f_len = torch.from_numpy(fl.lengths).to(dev)
    f_order = torch.argsort(f_len, descending=True).to(torch.int32)
    flat = {}
    for tml in (-3.0, -5.0):
        ms = _time_launches(torch, lambda: dec.decode_launch(f_logits, f_len, f_order, beam_width=200, token_min_logp=tml), 2, 1)
        flat[str(tml)] = {"ms": _max_over_ranks(ms, dev, world), "utt_per_s": world * 128 / ms * 1e3}
    clocks = sampler.stop(t0c, time.time()) if rank == 0 else None
    if rank == 0:
        frames = int(wl.lengths.sum())
        peak, src = _peak()
        ms5 = sweep["-5.0"]["ms"]
        line = _common("decoded utterances/sec (conversation-shaped, T<=1499, beam=200, 5-gram LM, token_min_logp sweep)",
                       "utterances/s", world, args.steps, args.warmup, ms5, "weak")
        line.update({
            "value": sweep["-5.0"]["utt_per_s"], "dtype": "f32 log-probs / f64 beam scores",
            "config": {"workload": "BASELINE configs[2]: conversation-shaped utterances (uniform 5-30 s, T<=1499), peaky logits, "
                                   "beam=200, beam_prune_logp=-10, 5-gram ARPA LM; value at token_min_logp=-5",
                       "utterances_per_gpu_per_step": n, "l2": "logits of one step: %.0f MB per GPU (> L2)" % (wl.logits.nbytes / 1e6)},
            "token_min_logp_sweep": sweep, "flat_logits_128_utterances": flat,
            "audio_s_per_s": world * float(synth.audio_seconds(wl.lengths).sum()) / ms5 * 1e3,
            "e2e": {"value": world * n / e2e_ms * 1e3, "unit": "utterances/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(frames * 46 * 4 + n * 8), "d2h_bytes_per_step": int(sum(map(len, wl.references)) * 4 + 8 * n),
                    "input": "pinned padded host logits -> decode_batch -> cer/wer"},
            "gpu_launches": args.steps, "roofline": {"bound": "hbm", "kernel": "beam_search_kernel<256,256,640>",
                                                     "achieved": frames * 46 * 4 / ms5 / 1e6, "peak": peak, "unit": "GB/s",
                                                     "frac": frames * 46 * 4 / ms5 / 1e6 / peak, "traffic": None, "peak_source": src,
                                                     "note": "logit bytes only (the LM probe counts need the oracle's cache-miss counters: see config 2)"},
            "parity_gate": gate, "quality": {"cer": cer_v, "wer": wer_v}, "clocks": clocks, "cpu_baseline": None,
        })
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ config 4
def run_config4(args, rank, world, local_rank, clock_sampler):
    import synth
    from coral_b200 import metrics
    from coral_b200.sharded import rates_from_totals, reduce_counts, shard_indices
    from coral_b200.textio import encode_utf32
    from coral_b200.validation import validation_scores

    N = 200_000 if args.utts == 8192 else args.utts
    rng = np.random.default_rng(4242)
    words = synth.make_word_list(50_000)
    idx = rng.integers(0, len(words), size=(N, 14))
    lens = rng.integers(3, 15, size=N)
    refs_all = [" ".join(words[j] for j in idx[i, : lens[i]]) for i in range(N)]
    hyps_all = [synth.corrupt_text(r, rng, 0.07) for r in refs_all]
    # CPU baseline (N = 1): the oracle's per-pair jiwer/rapidfuzz restatement on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import edit as oe

        s = 4000
        t0 = time.perf_counter()
        oe.cer(hyps_all[:s], refs_all[:s]); oe.wer(hyps_all[:s], refs_all[:s])
        cpu = {"value": s / (time.perf_counter() - t0), "unit": "pairs/s", "cores": 1, "kind": "port",
               "sample": f"first {s} pairs, single process (the reference loops over pairs in Python, R:src/coral/metrics.py:26-33)"}
    torch, dist, dev = _setup(local_rank, world)
    ref_len = np.array([len(r) for r in refs_all])
    mine = shard_indices(ref_len, rank, world)
    refs = [refs_all[i] for i in mine]
    hyps = [hyps_all[i] for i in mine]
    n = len(refs)
    r_cps, r_off = encode_utf32(refs)
    h_cps, h_off = encode_utf32(hyps)
    d = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).to(dev)
    d_r, d_ro, d_h, d_ho = d(r_cps), d(r_off), d(h_cps), d(h_off)
    max_len = int(max(np.diff(r_off).max(), np.diff(h_off).max()))
    cells = float((np.diff(r_off).astype(np.float64) * np.diff(h_off)).sum())
    kern = []

    def step_device():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cc, _ = metrics.edit_counts_device(d_r, d_ro, d_h, d_ho, n, 1, max_len, dev)
        e1.record()
        wc, _ = metrics.edit_counts_device(d_r, d_ro, d_h, d_ho, n, 2, max_len, dev)
        keep = (cc[:, :3].sum(dim=1).double() / (cc.sum(dim=1).double())) < 0.6
        tot = torch.stack([cc.sum(dim=0, dtype=torch.int64), wc.sum(dim=0, dtype=torch.int64)])
        if world > 1:
            dist.all_reduce(tot)
        kern.append((e0, e1))
        return tot, keep

    sampler = clock_sampler(local_rank)
    if rank == 0:
        sampler.start()
    t0c = time.time()
    dev_ms = _max_over_ranks(_time_launches(torch, step_device, args.steps, args.warmup), dev, world)
    chars_ms = float(np.median([a.elapsed_time(b) for a, b in kern[-args.steps:]]))

    shard_sizes = [len(shard_indices(ref_len, r, world)) for r in range(world)]

    def e2e():
        vs = validation_scores(hyps, refs, max_cer=0.6)
        totals = reduce_counts(vs.char_counts, vs.word_counts)
        cers, wers = rates_from_totals(totals)
        keep = torch.from_numpy(vs.keep).to(dev)
        if world > 1:  # the keep-mask of every shard, gathered (fixed-width, padded to the largest shard)
            sizes = shard_sizes
            pad = torch.zeros(max(sizes), dtype=torch.bool, device=dev)
            pad[: len(keep)] = keep
            out = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(out, pad)
            keep = torch.cat([o[:s] for o, s in zip(out, sizes)])
        return cers[0], wers[0], int(keep.sum().item()), vs

    for _ in range(2):
        e2e()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cer_v, wer_v, kept, vs = e2e()
    torch.cuda.synchronize()
    e2e_ms = _max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3, dev, world)
    clocks = sampler.stop(t0c, time.time()) if rank == 0 else None
    # parity gate: a 2000-pair subsample of this rank's shard against the oracle, bit-exact
    from oracle import edit as oe

    sub = np.random.default_rng(1).choice(n, size=min(2000, n), replace=False)
    for i in sub:
        assert tuple(vs.char_counts[i]) == oe.char_counts(refs[i], hyps[i]) and tuple(vs.word_counts[i]) == oe.word_counts(refs[i], hyps[i])
    if rank == 0:
        peak, src = _peak()
        alg = float(r_cps.nbytes + h_cps.nbytes + 16 * n)
        line = _common("dataset_validation scoring: (reference, hypothesis) pairs/sec, per-sample CER/WER + keep-mask + aggregate",
                       "pairs/s", world, args.steps, args.warmup, dev_ms, "strong")
        line.update({
            "value": N / dev_ms * 1e3, "dtype": "u32 code points / int32 counts",
            "config": {"workload": "BASELINE configs[3]: 200k synthetic (reference, 7%-corrupted hypothesis) pairs, read-aloud text shape, "
                                   "per-sample CER (normalised) + keep-mask cer<0.6 + aggregate CER/WER; pairs split over the ranks",
                       "pairs_total": N, "pairs_this_rank": n, "l2": "strings of one rank: %.1f MB (fits L2; scoring is ALU-bound)" % (alg / 1e6)},
            "e2e": {"value": N / e2e_ms * 1e3, "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(alg), "d2h_bytes_per_step": int(2 * n * 20),
                    "input": "Python lists of str -> validation_scores (UTF-32 encode, H2D, both kernels, D2H) -> all-reduce -> keep-mask all_gather"},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "kernel": "edit_bitpar_kernel(chars)", "achieved": alg / chars_ms / 1e6, "peak": peak,
                         "unit": "GB/s", "frac": alg / chars_ms / 1e6 / peak, "traffic": None, "peak_source": src,
                         "kernel_ms_per_launch": chars_ms, "GCUPS": cells / chars_ms / 1e6,
                         "note": "integer-ALU / latency bound (SURVEY 8d): GCUPS is the work rate, the HBM fraction is tiny by construction"},
            "quality": {"cer": cer_v, "wer": wer_v, "kept": kept, "kept_frac": kept / N},
            "parity_gate": {"per_pair_counts_bit_exact": True, "sample": int(len(sub))}, "cpu_baseline": cpu, "clocks": clocks,
        })
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ config 5
def run_config5(args, rank, world, local_rank, clock_sampler):
    import synth
    from coral_b200.decoder import build_ctcdecoder

    torch, dist, dev = _setup(local_rank, world)
    n = args.utts if args.utts != 8192 else 2048
    cache = os.environ.get("CORAL_B200_CACHE", "/tmp/coral_b200_cache")
    sampler = clock_sampler(local_rank)
    if rank == 0:
        sampler.start()
    t0c = time.time()
    grid = {}
    gate = None
    for order in (3, 4, 5, 6):
        wl = synth.build_workload(cache, n, order=order, kind="peaky", shape="read_aloud", name=f"sweep{rank}")
        dec = build_ctcdecoder(wl.labels, wl.arpa_path)
        if rank == 0 and order == 5:
            gate = _oracle_gate(wl, dec, dict(beam_width=128), n=8)
        d_logits = torch.from_numpy(wl.logits).to(dev)
        d_len = torch.from_numpy(wl.lengths).to(dev)
        d_order = torch.argsort(d_len, descending=True).to(torch.int32)
        for beam in (16, 32, 64, 128, 256, 512):
            ms = _time_launches(torch, lambda: dec.decode_launch(d_logits, d_len, d_order, beam_width=beam), args.steps, args.warmup)
            ms = _max_over_ranks(ms, dev, world)
            grid[f"order{order}_beam{beam}"] = {"ms": ms, "utt_per_s": world * n / ms * 1e3}
        frames = int(wl.lengths.sum())
        audio = float(synth.audio_seconds(wl.lengths).sum())
        del dec, d_logits
    clocks = sampler.stop(t0c, time.time()) if rank == 0 else None
    if rank == 0:
        peak, src = _peak()
        ms = grid["order5_beam128"]["ms"]
        line = _common("decoded utterances/sec (LM order 3-6 x beam 16-512 sweep; value at order 5, beam 128)", "utterances/s",
                       world, args.steps, args.warmup, ms, "weak")
        line.update({
            "value": grid["order5_beam128"]["utt_per_s"], "dtype": "f32 log-probs / f64 beam scores",
            "config": {"workload": "BASELINE configs[4]: 3/4/5/6-gram LM x beam 16-512 on read-aloud-shaped peaky logits",
                       "utterances_per_gpu_per_step": n, "l2": "logits of one step: %.0f MB per GPU (> L2)" % (n * 499 * 46 * 4 / 1e6)},
            "grid": grid, "audio_s_per_s": world * audio / ms * 1e3,
            "e2e": None, "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "kernel": "beam_search_kernel<128,128,320>", "achieved": frames * 46 * 4 / ms / 1e6,
                         "peak": peak, "unit": "GB/s", "frac": frames * 46 * 4 / ms / 1e6 / peak, "traffic": None, "peak_source": src,
                         "note": "logit bytes only; latency-bound kernel (see config 2)"},
            "parity_gate": gate, "clocks": clocks, "cpu_baseline": None,
        })
        print(json.dumps(line), flush=True)


def run(args, rank, world, local_rank, clock_sampler):
    import torch.distributed as dist

    {3: run_config3, 4: run_config4, 5: run_config5}[args.config](args, rank, world, local_rank, clock_sampler)
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()
