#!/usr/bin/env python
"""Benchmark of the hot path: batched CTC prefix beam search (beam 100, 5-gram LM shallow
fusion) + WER/CER scoring, on synthetic logits of the shape BASELINE.json names.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...           # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch: decode every utterance of the shard
(auto input detection + beam search kernel), turn the winning token rows into code points
on the device, score them against the references (chars + words edit-count kernels), sum
the counts and -- for N > 1 -- all-reduce them over NCCL. Workload = BASELINE.json
configs[1]: read-aloud-shaped utterances (T in [24, 499], mean ~290 frames), trained-model-
like ("peaky") logits, V = 46, beam 100, beam_prune_logp -10, token_min_logp -5, 5-gram ARPA
LM over a synthetic Danish-charset corpus. Weak scaling: every rank owns a full shard.

`value`   utterances/s with the inputs resident in HBM (CUDA events, max over ranks)
`e2e`     the same metric through the public API with HOST inputs, every step: pinned logits
          pulled over PCIe by the decode -> transcripts to the host as Python strings ->
          cer()/wer() (references H2D, counts D2H). `e2e.value` drives the steps through
          decoder.decode_batches (two batches ahead, as a dataset loop does);
          `e2e.one_call_at_a_time` is decode_batch + cer/wer with nothing overlapped
`roofline` for the beam-search kernel: algorithmic bytes / its CUDA-event time
`cpu_baseline` the oracle (a port of pyctcdecode+KenLM+jiwer, which are not installable
          here) timed on this box's host cores on a bounded sample of the same workload
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decoded utterances/sec (prefix beam search beam=100 + 5-gram LM fusion + WER/CER scoring)"
UNIT = "utterances/s"
CACHE = os.environ.get("CORAL_B200_CACHE", os.path.join(tempfile.gettempdir(), "coral_b200_cache"))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=8192, help="utterances per GPU per step")
    ap.add_argument("--beam", type=int, default=100)
    ap.add_argument("--order", type=int, default=5)
    ap.add_argument("--kind", default="peaky", choices=["peaky", "flat"])
    ap.add_argument("--shape", default="read_aloud", choices=["read_aloud", "conversation"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="utterances in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = the headline (default), 3 / 4 / 5 see bench_configs.py")
    return ap.parse_args()


# --------------------------------------------------------------------------- CPU side
_ORACLE = {}


def _oracle_init(labels, arpa):
    from oracle.beam import build_ctcdecoder

    _ORACLE["dec"] = build_ctcdecoder(labels, arpa)


def _oracle_decode(args):
    logits, beam = args
    dec = _ORACLE["dec"]
    s0 = dict(dec.stats)
    text = dec.decode_beams(logits, beam_width=beam)[0][0]
    d = {k: dec.stats[k] - s0[k] for k in s0}
    return text, d


def cpu_reference_pass(wl, idx, beam, n_procs, pool=None):
    """The reference's CPU path on utterances ``idx``: pyctcdecode-style beam search with the
    KenLM-semantics scorer, fork pool over utterances like
    HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:374-406, then cer()/wer()
    one pair at a time like R:src/coral/metrics.py:26-33. Returns (seconds, stats, hyps)."""
    from oracle import edit as oracle_edit

    items = [(wl.logits[u, : wl.lengths[u]], beam) for u in idx]
    t0 = time.perf_counter()
    if pool is None:
        out = [_oracle_decode(it) for it in items]
    else:
        out = pool.map(_oracle_decode, items, chunksize=max(1, len(items) // (4 * n_procs)))
    hyps = [o[0] for o in out]
    refs = [wl.references[u] for u in idx]
    c = oracle_edit.cer(hyps, refs)
    w = oracle_edit.wer(hyps, refs)
    dt = time.perf_counter() - t0
    stats = {}
    for _, d in out:
        for k, v in d.items():
            stats[k] = stats.get(k, 0) + v
    stats["cer"], stats["wer"] = c, w
    return dt, stats, hyps


def make_pool(wl, n_procs):
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    return ctx.Pool(n_procs, initializer=_oracle_init, initargs=(wl.labels, wl.arpa_path))


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions. In-process NVML (two
    cheap queries per sample); `nvidia-smi -lms` as the fallback. The period is a
    compromise: frequent polling of the driver measurably slows the kernels being timed."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int, period_s: float = 0.05):
        self.rows = []  # (time, sm_mhz, max_mhz, reasons bitmask)
        self.gpu = gpu_index
        self.period = float(os.environ.get("CORAL_BENCH_CLOCK_PERIOD_MS", period_s * 1e3)) / 1e3
        self.proc = None
        self._stop = threading.Event()
        self._thread = None

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        if self.period <= 0:
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            rs = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            rs = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.time(), sm, mx, rs))
                    except Exception:
                        pass
                    self._stop.wait(self.period)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            self.source = "nvml"
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                 str(max(20, int(self.period * 1e3))), "-i", str(self._physical_index())],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"

            def read():
                for line in self.proc.stdout:
                    p = [x.strip() for x in line.split(",")]
                    try:
                        rs = sum(bit for (_, bit), v in zip(self.REASONS, p[3:7]) if v.lower().startswith("active"))
                        self.rows.append((time.time(), float(p[1]), float(p[2]), rs))
                    except Exception:
                        continue

            threading.Thread(target=read, daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self, t_from=0.0, t_to=float("inf")):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r for r in self.rows if t_from <= r[0] <= t_to]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        mask = 0
        for r in rows:
            mask |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "reasons": sorted(name for name, bit in self.REASONS if mask & bit), "samples": len(rows),
                "source": getattr(self, "source", None), "period_ms": self.period * 1e3}


# ------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    import synth

    n_procs = os.cpu_count() or 1
    sample = args.cpu_sample or 16 * n_procs
    wl = synth.build_workload(CACHE, sample, order=args.order, kind=args.kind, shape=args.shape, name="eval0")
    idx = list(range(sample))
    pool = make_pool(wl, n_procs)
    try:
        for _ in range(max(args.warmup, 1)):
            cpu_reference_pass(wl, idx[: max(n_procs, sample // 4)], args.beam, n_procs, pool)
        times = []
        for _ in range(args.steps):
            dt, stats, _ = cpu_reference_pass(wl, idx, args.beam, n_procs, pool)
            times.append(dt)
    finally:
        pool.close()
        pool.join()
    total = sum(times)
    value = sample * args.steps / total
    audio = float(synth.audio_seconds(wl.lengths).sum())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 log-probs / f64 scores / int32 counts",
        "data": "synthetic",
        "config": dict(workload_config(args, sample, "bounded sample of the same workload; CPU only"),
                       sample_of={"utterances_per_step": sample, "of_utterances_per_step_in_the_gpu_arm": args.utts,
                                  "same_generator_and_seed": True}),
        "audio_s_per_s": audio * args.steps / total,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_procs, "kind": "port",
                         "sample": f"{sample} utterances per step, fork pool of {n_procs} processes over utterances "
                                   "(oracle port of pyctcdecode 0.5.0 + KenLM + jiwer; the real packages are not "
                                   "installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, utts, note):
    return {
        "workload": f"BASELINE configs[1]: {args.shape} {args.kind} logits [B,T<=499,46], beam={args.beam}, "
                    f"beam_prune_logp=-10, token_min_logp=-5, {args.order}-gram ARPA LM (50k words, 200k sentences), "
                    "alpha=0.5 beta=1.5 unk=-10, + CER/WER vs references",
        "utterances_per_gpu_per_step": utts, "beam_width": args.beam, "lm_order": args.order, "logits": args.kind,
        "l2": "inputs larger than L2 (logits of one step: %.0f MB per GPU)" % (utts * 499 * 46 * 4 / 1e6),
        "note": note,
    }


# ------------------------------------------------------------------------------ our arm
def oracle_counts_sample(wl, beam, n_procs, sample):
    """Oracle pass on the first ``sample`` utterances: (seconds, stats, hyps). Its cache-miss counts
    (n_score, probes, n_partial) define the algorithmic bytes of the beam kernel at EVERY N
     and its transcripts are the parity gate."""
    idx = list(range(sample))
    pool = make_pool(wl, n_procs)
    try:
        cpu_reference_pass(wl, idx[:n_procs], beam, n_procs, pool)  # warm the workers
        dt, stats, hyps = cpu_reference_pass(wl, idx, beam, n_procs, pool)
    finally:
        pool.close()
        pool.join()
    stats["sample"] = sample
    stats["hyps"] = hyps
    return dt, stats


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import synth
    from coral_b200 import metrics
    from coral_b200.decoder import build_ctcdecoder
    from coral_b200.textio import encode_utf32

    # ---- CPU side first (rank 0), before this process touches CUDA: the oracle sample gives the
    # parity gate, the algorithmic-byte counts and (N = 1 only) the reported CPU baseline
    cpu_baseline = None
    oracle_stats = None
    wl = synth.build_workload(CACHE, args.utts, order=args.order, kind=args.kind, shape=args.shape,
                              name=f"eval{rank}")
    if rank == 0 and not args.no_cpu_baseline:
        n_procs = os.cpu_count() or 1
        sample = args.cpu_sample or (16 * n_procs if world == 1 else 4 * n_procs)
        sample = min(sample, args.utts)
        dt, oracle_stats = oracle_counts_sample(wl, args.beam, n_procs, sample)
        if world == 1:
            _oracle_init(wl.labels, wl.arpa_path)
            t0 = time.perf_counter()
            seq_n = min(sample, 2 * n_procs)
            for u in range(seq_n):
                _oracle_decode((wl.logits[u, : wl.lengths[u]], args.beam))
            seq_rate = seq_n / (time.perf_counter() - t0)
            cpu_baseline = {
                "value": sample / dt, "unit": UNIT, "cores": n_procs, "kind": "port",
                "sample": f"first {sample} utterances of the workload, fork pool of {n_procs} processes over utterances "
                          f"(HF batch_decode regime) + cer/wer; sequential single-process regime (what evaluate() does): "
                          f"{seq_rate:.1f} utt/s; oracle port -- pyctcdecode/kenlm/jiwer are not installable here",
            }

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # each rank keeps to its share of the host cores (eight ranks on one socket otherwise fight
    # over the same cores for their string work)
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
    except Exception:
        pass

    dec = build_ctcdecoder(wl.labels, wl.arpa_path)
    B = args.utts
    Tm = wl.logits.shape[1]
    h_logits = torch.from_numpy(wl.logits).pin_memory()
    h_len = torch.from_numpy(wl.lengths).pin_memory()
    d_logits = h_logits.to(dev)
    d_len = h_len.to(dev)
    refs = wl.references
    r_cps, r_off = encode_utf32(refs)
    d_rcps = torch.from_numpy(r_cps.view(np.int32)).to(dev)
    d_roff = torch.from_numpy(r_off).to(dev)
    ref_max_len = int(np.diff(r_off).max())
    kern_ms = []

    def step_device(time_kernel=False):
        """Inputs resident in HBM; everything stays on the device."""
        if time_kernel:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d_order = torch.argsort(d_len, descending=True).to(torch.int32)
        outs = dec.decode_launch(d_logits, d_len, d_order, beam_width=args.beam, n_best=1,
                                 events=(e0, e1) if time_kernel else None)
        d_n, d_logit, d_comb, d_tok, d_lens, d_status = outs
        h_cps, h_off, h_max = dec.device_text(d_tok, d_lens)   # transcripts as UTF-32 on the device
        # sizes the edit kernel's on-chip buffers: one scalar read back (a sync inside the step)
        max_len = max(ref_max_len, int(h_max.item()))
        cc, _ = metrics.edit_counts_device(d_rcps, d_roff, h_cps, h_off, B, 1, max_len, dev)
        wc, _ = metrics.edit_counts_device(d_rcps, d_roff, h_cps, h_off, B, 2, max_len, dev)
        totals = torch.stack([cc.sum(dim=0, dtype=torch.int64), wc.sum(dim=0, dtype=torch.int64)])
        if world > 1:
            dist.all_reduce(totals)
        if time_kernel:
            kern_ms.append((e0, e1))
        return totals, d_status

    def step_e2e(inp=None, lengths=h_len):
        """Public API with HOST inputs: the logits reach the GPU inside decode_batch (pinned host
        memory read in place), transcripts come back as strings, cer()/wer() score them."""
        hyps = dec.decode_batch(None, h_logits if inp is None else inp, beam_width=args.beam, lengths=lengths)
        if world > 1:
            from coral_b200.sharded import sharded_error_rates

            r = sharded_error_rates(hyps, refs)
            return hyps, r["cer"], r["wer"]
        return hyps, metrics.cer(hyps, refs), metrics.wer(hyps, refs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate on the CPU sample before any timing is reported
    hyps, cer_v, wer_v = step_e2e()
    parity = None
    if oracle_stats is not None:
        s = oracle_stats["sample"]
        same = list(hyps[:s]) == oracle_stats["hyps"]
        from oracle import edit as oracle_edit

        same_counts = (metrics.cer(hyps[:s], refs[:s]) == oracle_edit.cer(hyps[:s], refs[:s]) and
                       metrics.wer(hyps[:s], refs[:s]) == oracle_edit.wer(hyps[:s], refs[:s]))
        parity = {"transcripts_identical": bool(same), "cer_wer_bit_exact": bool(same_counts), "sample": s}
        if not (same and same_counts):
            raise SystemExit(f"PARITY FAILURE against the oracle on the CPU sample: {parity}")

    # ---- timed: device-resident
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # started before the warm-up: nvidia-smi needs ~100 ms to come up
    for _ in range(max(args.warmup, 3)):
        totals, d_status = step_device()
    barrier()
    t_clock0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        totals, d_status = step_device(time_kernel=True)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    beam_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ms]))
    if rank == 0:
        print("beam kernel ms per step: " + " ".join(f"{a.elapsed_time(b):.2f}" for a, b in kern_ms), file=sys.stderr)
    assert int(d_status.sum().item()) == 0
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- timed: end to end through the public API, ALL --steps
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hyps, cer_v, wer_v = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    # clocks sampled during the two timed regions (device-resident steps and end-to-end steps)
    clocks = sampler.stop(t_clock0, time.time()) if rank == 0 else None
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    # ---- the same steps through decode_batches: batch k + 1 is decoding while the host turns batch k
    # into strings and scores it (the dataset loop of evaluate() / add_validations)
    def score(hy, labels=None):
        labels = refs if labels is None else labels
        if world > 1:
            from coral_b200.sharded import sharded_error_rates

            r = sharded_error_rates(hy, labels)
            return r["cer"], r["wer"]
        return metrics.cer(hy, labels), metrics.wer(hy, labels)

    def run_pipelined(n):
        last = None
        for hy in dec.decode_batches(((h_logits, h_len) for _ in range(n)), beam_width=args.beam):
            last = (hy, *score(hy))
        return last

    hy_p, cer_p, wer_p = run_pipelined(3)
    assert list(hy_p) == list(hyps) and (cer_p, wer_p) == (cer_v, wer_v)
    barrier()
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pipe_s = float(t.item())

    # ---- where an end-to-end step spends its time (separate, untimed-for-the-headline pass with a
    # synchronisation after every phase; max over ranks)
    def e2e_phases(n=5):
        """Per-phase wall clock of an end-to-end step with a synchronisation after every phase:
        median over n repetitions, max over ranks."""
        acc = {}

        def mark(name, t_prev):
            torch.cuda.synchronize()
            now = time.perf_counter()
            acc.setdefault(name, []).append(now - t_prev)
            return now

        for _ in range(n):
            barrier()
            t_ = time.perf_counter()
            outs = dec.decode_padded(h_logits, h_len, beam_width=args.beam, n_best=1, to_host=False)
            t_ = mark("decode (launch prep + kernel pulling the pinned host logits)", t_)
            d_text = dec.device_text(outs[3], outs[4])
            bad = int(outs[5].abs().max().item())
            t_ = mark("transcripts -> UTF-32 on the device + status read-back", t_)
            texts = dec._texts_from_device(*d_text)
            t_ = mark("D2H of the text + Python strings", t_)
            metrics._to_device(refs, dev)
            t_ = mark("references: UTF-32 encode + H2D", t_)
            if world > 1:
                from coral_b200.sharded import sharded_error_rates

                sharded_error_rates(texts, refs)
            else:
                metrics.cer(texts, refs), metrics.wer(texts, refs)
            t_ = mark("cer + wer: both edit kernels, one read-back" + (", all-reduce" if world > 1 else ""), t_)
            assert bad == 0
        v = torch.tensor([float(np.median(acc[k])) * 1e3 for k in acc], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return {k: round(float(x), 3) for k, x in zip(acc, v.cpu().tolist())}

    phases = e2e_phases()

    # ---- the reference's real call shape: a list of [T_i, V] numpy arrays (what
    # HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:371 hands to decode_beams_batch),
    # pageable host memory, packed into the pinned ragged buffer by host threads
    lst = [wl.logits[u, : wl.lengths[u]] for u in range(B)]
    for _ in range(2):
        hy = dec.decode_batch(None, lst, beam_width=args.beam)
    assert list(hy) == list(hyps)
    n_list = max(1, min(args.steps, 5))

    def timed_host(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def list_one_at_a_time():
        for _ in range(n_list):
            score(dec.decode_batch(None, lst, beam_width=args.beam))

    def list_pipelined(n=n_list):
        for hy in dec.decode_batches((lst for _ in range(n)), beam_width=args.beam):
            score(hy)

    list_s = timed_host(list_one_at_a_time)
    list_pipelined(2)
    list_pipe_s = timed_host(list_pipelined)

    # ---- evaluate()-shaped end to end (R:src/coral/evaluate.py:56-84): decode -> text normaliser on
    # every transcript (numerals, lower, NFKC, conversion dict, characters_to_keep) -> cer / wer
    from coral_b200.normalise import DEFAULT_CONVERSION_DICT, TextNormaliser

    norm = TextNormaliser("abcdefghijklmnopqrstuvwxyzæøå0123456789éü", DEFAULT_CONVERSION_DICT,
                          lower_case=True, convert_numerals=True)   # R:config/evaluation.yaml:14
    refs_norm = norm(refs)

    def step_evaluate():
        hy = norm(dec.decode_batch(None, h_logits, beam_width=args.beam, lengths=h_len))
        if world > 1:
            from coral_b200.sharded import sharded_error_rates

            r = sharded_error_rates(hy, refs_norm)
            return r["cer"], r["wer"]
        return metrics.cer(hy, refs_norm), metrics.wer(hy, refs_norm)

    for _ in range(2):
        step_evaluate()
    n_eval = max(1, min(args.steps, 5))

    def eval_one_at_a_time():
        for _ in range(n_eval):
            step_evaluate()

    def eval_pipelined(n=n_eval):
        for hy in dec.decode_batches(((h_logits, h_len) for _ in range(n)), beam_width=args.beam):
            score(norm(hy), refs_norm)

    eval_s = timed_host(eval_one_at_a_time)
    eval_pipelined(2)
    eval_pipe_s = timed_host(eval_pipelined)

    # ---- work counters (one extra untimed launch with stats)
    st = dec.decode_padded(d_logits, d_len, beam_width=args.beam, n_best=1, collect_stats=True).stats

    # ---- the other kernels of the path, timed alone on the same inputs (rank 0; not part of `value`)
    other = None
    if rank == 0:
        from coral_b200.greedy import greedy_decode_device

        def timed(fn, n=20):
            """Mean device time of n back-to-back launches (events around the whole train: a host
            sync per launch would add the Python launch overhead to a ~100 us kernel)."""
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        g_ms = timed(lambda: greedy_decode_device(d_logits, d_len, blank_id=45))
        fr = int(wl.lengths.sum())
        g_bytes = fr * 46 * 4 + fr * 4  # every valid logit read once + the collapsed ids' upper bound
        o0 = dec.decode_launch(d_logits, d_len, None, beam_width=args.beam, n_best=1)
        h_cps0, h_off0, h_max0 = dec.device_text(o0[3], o0[4])
        ml = max(ref_max_len, int(h_max0.item()))
        e_ms = timed(lambda: metrics.edit_counts_device(d_rcps, d_roff, h_cps0, h_off0, B, 1, ml, dev))
        hl = (h_off0[1:] - h_off0[:-1]).double()
        cells = float((torch.from_numpy(np.diff(r_off)).to(dev).double() * hl).sum().item())
        other = {
            "ctc_greedy (argmax + collapse)": {
                "ms": g_ms, "algorithmic_bytes": g_bytes, "GB_per_s": g_bytes / g_ms / 1e6, "bound": "hbm",
                "note": "greedy decode of the same ragged logits (configs[0] / config 4 path); inputs 425 MB > L2"},
            "edit_bitpar_kernel(chars)": {
                "ms": e_ms, "cells": cells, "GCUPS": cells / e_ms / 1e6, "bound": "integer ALU (bit-parallel recurrence), memory latency"},
        }

    if rank == 0:
        frames = int(wl.lengths.sum())
        audio = float(synth.audio_seconds(wl.lengths).sum())
        hyp_chars = int(sum(len(h) for h in hyps))
        # algorithmic bytes of the beam kernel: logits once + LM slots + output.
        # ONE definition at every N: the oracle's distinct (cache-miss) counts on rank 0's CPU sample.
        alg_logits = frames * 46 * 4
        if oracle_stats is not None:
            s = oracle_stats["sample"]
            scale = B / s
            n_score, probes, n_partial = (oracle_stats[k] * scale for k in ("n_score", "probes", "n_partial"))
            lm_src = f"oracle cache-miss counts on rank 0's first {s} utterances, scaled to {B}"
        else:
            n_score, probes, n_partial = float(st[1]), float(st[2]), float(st[4])
            lm_src = "device counters (--no-cpu-baseline: no oracle sample in this run)"
        alg_lm = probes * 16 + n_score * 16 + n_partial * 8
        alg_out = hyp_chars + 16 * B
        alg_bytes = alg_logits + alg_lm + alg_out
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        if other:
            g = other["ctc_greedy (argmax + collapse)"]
            g["frac_of_hbm_peak"] = g["GB_per_s"] / peak
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "beam_kernel_traffic.json"))).get("dram_bytes_per_launch") if B == 8192 else None
        except Exception:
            pass
        achieved = alg_bytes / (beam_ms * 1e-3) / 1e9
        e2e_value = world * B * args.steps / pipe_s
        ref_bytes = r_cps.nbytes + r_off.nbytes
        line = {
            "metric": METRIC, "value": world * B * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 log-probs / f64 beam scores / int32 edit counts", "data": "synthetic",
            "config": workload_config(args, B, "value: inputs resident in HBM; e2e: host logits through decode_batch + cer/wer"),
            "audio_s_per_s": world * audio * args.steps / (dev_ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    # pinned host logits are pulled by the kernel itself: the valid frames cross PCIe once
                    "h2d_bytes_per_step": int(frames * 46 * 4 + 2 * B * 4 + ref_bytes),
                    "d2h_bytes_per_step": int(hyp_chars * 4 + 8 * (B + 1) + 2 * B * 20 + 8),
                    "audio_s_per_s": world * audio * args.steps / pipe_s, "steps": args.steps,
                    "ms_per_step": 1e3 * pipe_s / args.steps,
                    "api": "for hyps in decoder.decode_batches(batches): cer(hyps, refs); wer(hyps, refs) -- batch k + 1 "
                           "decodes while the host builds the strings of batch k and scores them; every step's logits "
                           "cross PCIe and every step's strings and error rates come back to the host inside the timed region",
                    "input": "pinned padded [B, T_max, V] host tensor + lengths (decode_batch extension)",
                    "one_call_at_a_time": {"value": world * B * args.steps / e2e_s, "unit": UNIT, "steps": args.steps,
                                           "ms_per_step": 1e3 * e2e_s / args.steps,
                                           "api": "hyps = decoder.decode_batch(None, logits, lengths=...); cer(hyps, refs); "
                                                  "wer(hyps, refs) -- nothing overlapped across steps"},
                    "phases_ms": phases,
                    "evaluate_shaped": {"value": world * B * n_eval / eval_pipe_s, "unit": UNIT, "steps": n_eval,
                                        "ms_per_step": 1e3 * eval_pipe_s / n_eval,
                                        "one_call_at_a_time": world * B * n_eval / eval_s,
                                        "path": "decode_batch -> C++ text normaliser (numerals, lower, NFKC, conversion dict, "
                                                "characters_to_keep; host threads) -> cer/wer, as evaluate() does it "
                                                "(R:src/coral/evaluate.py:56-84)"},
                    "list_input": {"value": world * B * n_list / list_pipe_s, "unit": UNIT, "steps": n_list,
                                   "ms_per_step": 1e3 * list_pipe_s / n_list,
                                   "one_call_at_a_time": world * B * n_list / list_s,
                                   "input": "list of B pageable [T_i, V] numpy arrays -> decode_batch(None, list) + cer/wer "
                                            "(the call shape of HF Wav2Vec2ProcessorWithLM.batch_decode -> decode_beams_batch)"}},
            "gpu_launches": 8 * args.steps,
            "kernels_per_step": ["classify_heavy_kernel", "beam_search_kernel (lean)", "beam_search_kernel (heavy frames)",
                                 "text_count_kernel", "text_scan_kernel", "text_write_kernel",
                                 "edit_bitpar_kernel(chars)", "edit_bitpar_kernel(words)"],
            "roofline": {"bound": "hbm", "kernel": "beam_search_kernel<128,104,208> (beam widths <= 104; <128,128,320> up to 128)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_breakdown": {"logits": alg_logits, "lm_slots": alg_lm, "output": alg_out, "lm_counts_from": lm_src},
                         "kernel_ms_per_launch": beam_ms,
                         "note": "latency-bound by construction (SURVEY 7.6): T sequential frames per utterance; "
                                 "see beam_extensions_per_s and profiles/ for stall reasons"},
            "beam_extensions_per_s": float(st[0]) / (beam_ms * 1e-3),
            "device_counters_per_step": {"beam_extensions": int(st[0]), "lm_word_scorings": int(st[1]),
                                         "ngram_probes": int(st[2]), "frames": int(st[3]), "lexicon_probes": int(st[4])},
            "other_kernels": other,
            "cpu_baseline": cpu_baseline,
            "parity_gate": parity,
            "quality": {"cer": cer_v, "wer": wer_v},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config != 2:
        import bench_configs

        bench_configs.run(args, rank, world, local_rank, ClockSampler)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
