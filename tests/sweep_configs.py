"""BASELINE configs 3 and 5: LM order x beam width sweep and the long-utterance / token-threshold
sweep. Kernel-only utt/s (CUDA events, min of 3) plus a transcript parity check against the oracle
on a couple of utterances per cell. Prints one JSON line per cell."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/: it uses the oracle as checker)
sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from coral_b200.decoder import build_ctcdecoder
from oracle.beam import build_ctcdecoder as oracle_build

cache = os.path.join(tempfile.gettempdir(), "coral_b200_cache")
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def run_cell(dec, odec, wl, d_logits, d_len, beam, tmin, n_check=2, tag=None):
    d_order = torch.argsort(d_len, descending=True).to(torch.int32)
    for _ in range(1):
        dec.decode_launch(d_logits, d_len, d_order, beam_width=beam, token_min_logp=tmin)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        outs = dec.decode_launch(d_logits, d_len, d_order, beam_width=beam, token_min_logp=tmin, events=(e0, e1))
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    assert int(outs[5].sum().item()) == 0, "capacity status"
    # parity on the shortest utterances (the oracle is slow at wide beams)
    idx = np.argsort(wl.lengths)[len(wl.lengths) // 4:][:n_check]
    texts = dec.tokens_to_text(outs[3][:, 0, :].cpu().numpy(), outs[4][:, 0].cpu().numpy())
    ok = True
    for u in idx:
        ref = odec.decode_beams(wl.logits[u, : wl.lengths[u]], beam_width=beam, token_min_logp=tmin)[0][0]
        ok = ok and (ref == texts[u])
    ms = min(ts)
    rec = dict(tag, beam=beam, token_min_logp=tmin, utts=len(wl.lengths), ms=round(ms, 2),
               utt_per_s=round(len(wl.lengths) / ms * 1e3), audio_s_per_s=round(wl.audio_seconds / ms * 1e3),
               oracle_transcripts_identical=bool(ok))
    print(json.dumps(rec), flush=True)


if which in ("all", "config5"):
    for order in (3, 4, 5, 6):
        wl = synth.build_workload(cache, 2048, order=order, name="sweep")
        dec = build_ctcdecoder(wl.labels, wl.arpa_path)
        odec = oracle_build(wl.labels, wl.arpa_path)
        d_logits = torch.from_numpy(wl.logits).to(dev)
        d_len = torch.from_numpy(wl.lengths).to(dev)
        for beam in (16, 32, 64, 128, 256, 512):
            run_cell(dec, odec, wl, d_logits, d_len, beam, -5.0, tag=dict(config=5, lm_order=order, logits="peaky"))
        del dec, d_logits

if which in ("all", "config3"):
    for kind in ("peaky",):
        wl = synth.build_workload(cache, 1024, order=5, kind=kind, shape="conversation", name="conv")
        dec = build_ctcdecoder(wl.labels, wl.arpa_path)
        odec = oracle_build(wl.labels, wl.arpa_path)
        d_logits = torch.from_numpy(wl.logits).to(dev)
        d_len = torch.from_numpy(wl.lengths).to(dev)
        for tmin in (-3.0, -5.0, -7.0, -10.0, -20.0):
            run_cell(dec, odec, wl, d_logits, d_len, 200, tmin, n_check=1,
                     tag=dict(config=3, lm_order=5, logits=kind, T_max=int(wl.lengths.max())))

if which in ("all", "config3flat"):
    wl = synth.build_workload(cache, 256, order=5, kind="flat", shape="conversation", name="convflat")
    dec = build_ctcdecoder(wl.labels, wl.arpa_path)
    odec = oracle_build(wl.labels, wl.arpa_path)
    d_logits = torch.from_numpy(wl.logits).to(dev)
    d_len = torch.from_numpy(wl.lengths).to(dev)
    for tmin in (-3.0, -5.0):
        run_cell(dec, odec, wl, d_logits, d_len, 200, tmin, n_check=1,
                 tag=dict(config=3, lm_order=5, logits="flat", T_max=int(wl.lengths.max())))

if which in ("all", "config2"):
    # the headline workload presented as one batch, in the reference's batches of 16
    # (R:config/evaluation.yaml:20), and its flat-logit stress variant
    for kind in ("peaky", "flat"):
        n = 8192 if kind == "peaky" else 1776
        wl = synth.build_workload(cache, n, order=5, kind=kind, name="eval0")
        dec = build_ctcdecoder(wl.labels, wl.arpa_path)
        odec = oracle_build(wl.labels, wl.arpa_path)
        d_logits = torch.from_numpy(wl.logits).to(dev)
        d_len = torch.from_numpy(wl.lengths).to(dev)
        run_cell(dec, odec, wl, d_logits, d_len, 100, -5.0, n_check=2,
                 tag=dict(config=2, lm_order=5, logits=kind, presentation="one batch"))
        if kind == "peaky":
            nb = 1024
            for _ in range(2):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for a0 in range(0, nb, 16):
                    dec.decode_launch(d_logits[a0:a0 + 16], d_len[a0:a0 + 16], None, beam_width=100)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
            print(json.dumps(dict(config=2, lm_order=5, logits=kind, presentation="batches of 16, one stream, back to back",
                                  beam=100, utts=nb, ms=round(ms, 2), utt_per_s=round(nb / ms * 1e3))), flush=True)
            streams = [torch.cuda.Stream() for _ in range(16)]
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k, a0 in enumerate(range(0, nb, 16)):
                    with torch.cuda.stream(streams[k % len(streams)]):
                        dec.decode_launch(d_logits[a0:a0 + 16], d_len[a0:a0 + 16], None, beam_width=100)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) * 1e3
            print(json.dumps(dict(config=2, lm_order=5, logits=kind, presentation="batches of 16 round-robin over 16 streams (wall clock)",
                                  beam=100, utts=nb, ms=round(ms, 2), utt_per_s=round(nb / ms * 1e3))), flush=True)
        del dec, d_logits
