"""``BeamSearchDecoderCTC`` / ``build_ctcdecoder``: pyctcdecode's surface over the CUDA decoder.

Host-side mirror of UP:pyctcdecode 0.5.0 ``decoder.py``: same
names, argument meaning, defaults, return tuples and error behaviour, so that
``Wav2Vec2ProcessorWithLM`` (HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:
80-84, :143, :160-206, :365-367, :398-406, :565-572), the HF ASR pipeline
(HF:pipelines/automatic_speech_recognition.py:612-623) and CoRal
(R:src/coral/ngram.py:341-348) can use it unchanged. All decoding runs in
``coral_ctc_beam_decode`` on the GPU, a whole batch per launch; the ``pool`` argument is
accepted and ignored (no process pool is needed).
"""

from __future__ import annotations

import ctypes as C
import logging
import os

import numpy as np

from . import _lib
from .alphabet import (
    DEFAULT_ALPHA,
    DEFAULT_BEAM_WIDTH,
    DEFAULT_BETA,
    DEFAULT_HOTWORD_WEIGHT,
    DEFAULT_MIN_TOKEN_LOGP,
    DEFAULT_PRUNE_BEAMS,
    DEFAULT_PRUNE_LOGP,
    DEFAULT_SCORE_LM_BOUNDARY,
    DEFAULT_UNK_LOGP_OFFSET,
    Alphabet,
)
from .language_model import KenlmModel, LanguageModel, load_unigram_set_from_arpa
from .textio import encode_utf32

logger = logging.getLogger(__name__)

MAX_BEAM_WIDTH = 512
H2D_CHUNK = int(os.environ.get("CORAL_H2D_CHUNK", "512"))  # utterances per host-side packing chunk
# (large chunks: a launch with few utterances per CTA is dominated by its longest utterances)
# Host logits: "zerocopy" (default) = the kernel reads pinned host memory in place, valid frames
# only; "dma" = round 1's chunked cudaMemcpyAsync of the padded batch on a copy stream.
HOST_INPUT = os.environ.get("CORAL_HOST_INPUT", "zerocopy")


def _host_threads() -> int:
    """Threads for coral_host_pack_rows: this process's share of the cores it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, min(8, n // max(local, 1)))


def _invalidating(name):
    def method(self, *a, **kw):
        self._coral_dev = None  # the list no longer holds what was decoded
        return getattr(list, name)(self, *a, **kw)

    method.__name__ = name
    return method


class DecodedTexts(list):
    """``list[str]`` of transcripts that also remembers where their UTF-32 code points sit on the
    device (``_coral_dev`` = (cps int32 tensor, offsets int64 tensor, max_len, token)), so that
    ``coral_b200.metrics`` can score them without re-encoding and re-uploading the strings. Any
    in-place mutation drops the device copy; a new list built from it never had one."""

    __slots__ = ("_coral_dev",)
    _tokens = iter(range(1, 1 << 62))

    for _m in ("__setitem__", "__delitem__", "__iadd__", "__imul__", "append", "extend", "insert", "pop",
               "remove", "sort", "reverse", "clear"):
        locals()[_m] = _invalidating(_m)
    del _m


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("coral_b200 needs a CUDA device: there is no CPU path")
    return torch


class DecodedBatch:
    """Raw result of one batched launch (host numpy arrays)."""

    def __init__(self, n_beams, logit, comb, tokens, lens, status, stats=None, word_frames=None, word_counts=None):
        self.word_frames = word_frames   # int32 [B, n_best, max_words, 2] or None
        self.word_counts = word_counts   # int32 [B, n_best] or None
        self.n_beams = n_beams
        self.logit_score = logit
        self.lm_score = comb
        self.tokens = tokens
        self.lens = lens
        self.status = status
        self.stats = stats


class BeamSearchDecoderCTC:
    # pyctcdecode keeps language models in a class-level container keyed by a random id
    # so that fork pools share them; HF reads and writes attributes through it.
    model_container: dict[bytes, LanguageModel | None] = {}

    _ALPHABET_SERIALIZED_FILENAME = "alphabet.json"
    _LANGUAGE_MODEL_SERIALIZED_DIRECTORY = "language_model"

    def __init__(self, alphabet: Alphabet, language_model: LanguageModel | None = None,
                 device: int | None = None) -> None:
        self._alphabet = alphabet
        self._idx2vocab = {n: c for n, c in enumerate(self._alphabet.labels)}
        self._is_bpe = alphabet.is_bpe
        self._model_key = os.urandom(16)
        BeamSearchDecoderCTC.model_container[self._model_key] = language_model
        self._device = device
        self._h = None
        labels = self._alphabet.labels
        if len(labels) > 64:
            raise NotImplementedError("alphabets above 64 entries are not supported by this build")
        self._blank_id = labels.index("")
        self._space_id = labels.index(" ") if " " in labels else -1
        # id -> code point table for the fast token->string path (single-code-point labels)
        self._single_cp = all(len(c) <= 1 for c in labels)
        self._cp_table = np.array([ord(c) if len(c) == 1 else 0 for c in labels], dtype=np.uint32)
        self._latin1 = all(ord(ch) < 256 for c in labels for ch in c)

    # ------------------------------------------------------------------ plumbing
    @property
    def _language_model(self) -> LanguageModel | None:
        return BeamSearchDecoderCTC.model_container[self._model_key]

    def cleanup(self) -> None:
        BeamSearchDecoderCTC.model_container.pop(self._model_key, None)

    def reset_params(self, alpha=None, beta=None, unk_score_offset=None, lm_score_boundary=None) -> None:
        language_model = self._language_model
        if language_model is None:
            return
        if alpha is not None:
            language_model.alpha = alpha
        if beta is not None:
            language_model.beta = beta
        if unk_score_offset is not None:
            language_model.unk_score_offset = unk_score_offset
        if lm_score_boundary is not None:
            language_model.score_boundary = lm_score_boundary

    def _handle(self):
        if self._h is None:
            torch = _torch()
            lm = self._language_model
            if self._device is None:
                self._device = lm.kenlm_model.device if lm is not None else torch.cuda.current_device()
            lib = _lib.load()
            cps, off = encode_utf32(self._alphabet.labels)
            off32 = off.astype(np.int32)
            h = C.c_void_p()
            if lm is not None and lm.unigrams is not None:
                ucps, uoff = encode_utf32(lm.unigrams)
                n_uni = len(lm.unigrams)
            else:
                ucps, uoff, n_uni = np.zeros(1, np.uint32), np.zeros(1, np.int64), -1
            _lib.check(lib.coral_decoder_create(
                cps.ctypes.data, off32.ctypes.data, len(self._alphabet.labels), self._blank_id, self._space_id,
                lm.kenlm_model._h if lm is not None else None, ucps.ctypes.data, uoff.ctypes.data, n_uni,
                self._device, C.byref(h)))
            self._h = h
        lm = self._language_model
        if lm is not None:
            _lib.check(_lib.load().coral_decoder_set_params(
                self._h, float(lm.alpha), float(lm.beta), float(lm.unk_score_offset), int(bool(lm.score_boundary))))
        return self._h

    def __deepcopy__(self, memo):
        # HF's ProcessorMixin.to_dict() deep-copies its attributes; native handles are shared
        return self

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().coral_decoder_free(self._h)
                self._h = None
        except Exception:
            pass

    def _check_logits_dimension(self, logits) -> None:
        if len(logits.shape) != 2:
            raise ValueError(
                "Input logits have %s dimensions, but need 2: (time, vocabulary)" % len(logits.shape)
            )
        if logits.shape[-1] != len(self._idx2vocab):
            raise ValueError(
                "Input logits shape is %s, but vocabulary is size %s. "
                "Need logits of shape: (time, vocabulary)" % (logits.shape, len(self._idx2vocab))
            )

    # --------------------------------------------------------------- batched core
    def decode_padded(self, logits, lengths, beam_width: int = DEFAULT_BEAM_WIDTH,
                      beam_prune_logp: float = DEFAULT_PRUNE_LOGP, token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP,
                      prune_history: bool = False, n_best: int = 1, input_mode: int = 0,
                      collect_stats: bool = False, to_host: bool = True, word_frames: bool = False):
        """Decode a padded batch ``[B, T_max, V]`` (torch CUDA/CPU tensor or numpy) in one launch.

        ``lengths`` int [B]. Returns :class:`DecodedBatch` (numpy) or, with
        ``to_host=False``, the tuple of device tensors (the launch stays asynchronous).
        ``word_frames=True`` selects the kernel that also tracks pyctcdecode's text_frames.
        """
        torch = _torch()
        h = self._handle()
        dev = torch.device("cuda", self._device)
        if isinstance(logits, np.ndarray):
            logits = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32))
        if logits.dim() != 3:
            raise ValueError("Input logits have %s dimensions, but need 3: (batch, time, vocabulary)" % logits.dim())
        if logits.shape[-1] != len(self._idx2vocab):
            raise ValueError(
                "Input logits shape is %s, but vocabulary is size %s. "
                "Need logits of shape: (time, vocabulary)" % (tuple(logits.shape[1:]), len(self._idx2vocab)))
        if not 1 <= beam_width <= MAX_BEAM_WIDTH:
            raise ValueError(f"beam_width must be in [1, {MAX_BEAM_WIDTH}]")
        n_best = max(1, min(int(n_best), int(beam_width)))
        if isinstance(lengths, np.ndarray):
            lengths = torch.from_numpy(np.ascontiguousarray(lengths))
        elif not torch.is_tensor(lengths):
            lengths = torch.tensor(list(lengths))
        d_len = lengths.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        d_stats = torch.zeros(32, dtype=torch.int64, device=dev) if collect_stats else None
        B = logits.shape[0]
        if logits.device.type == "cpu" and logits.dtype == torch.float32 and logits.is_contiguous() and \
                logits.is_pinned() and HOST_INPUT == "zerocopy":
            # Pinned host logits are read IN PLACE by the kernel (unified addressing makes pinned
            # memory device-accessible): each valid frame crosses PCIe exactly once, padding rows
            # never move, and there is no staging copy to wait for.
            d_order = torch.argsort(d_len, descending=True).to(torch.int32)
            self._keepalive = (logits,)  # the host buffer must outlive the launch
            d_n, d_logit, d_comb, d_tok, d_lens, d_status, *wf = self._launch_raw(
                logits.data_ptr(), B, int(logits.shape[1]), int(logits.shape[2]), dev, d_len, d_order, beam_width,
                beam_prune_logp, token_min_logp, n_best, input_mode, d_stats, word_frames=word_frames,
                prune_history=prune_history)
        elif logits.device.type == "cpu" and HOST_INPUT == "zerocopy" and B > 0:
            # pageable host logits: the valid rows are packed into a pinned ragged staging buffer
            # by host threads (chunk by chunk, the kernel already running), then read in place
            lens_np = np.ascontiguousarray(lengths.cpu().numpy() if torch.is_tensor(lengths) else lengths).astype(np.int64)
            arr = logits.numpy() if logits.dtype == torch.float32 and logits.is_contiguous() else \
                np.ascontiguousarray(logits.to(torch.float32).numpy())
            base = arr.ctypes.data
            pitch = arr.shape[1] * arr.shape[2] * 4
            ptrs = base + np.arange(B, dtype=np.int64) * pitch
            self._keepalive = (arr,)
            d_n, d_logit, d_comb, d_tok, d_lens, d_status, *wf = self._decode_rows(
                ptrs, lens_np, int(arr.shape[1]), dev, beam_width, beam_prune_logp, token_min_logp, n_best,
                input_mode, d_stats, word_frames, prune_history)
        elif (logits.device.type == "cpu" and B >= 2 * H2D_CHUNK and logits.dtype == torch.float32
                and logits.is_contiguous() and logits.is_pinned()):
            # (CORAL_HOST_INPUT=dma) ONE launch over the whole batch, fed by a copy stream. The
            # copier delivers H2D_CHUNK utterances at a time and bumps a device counter after
            # each chunk; thread groups wait for their utterance's chunk (include/coral_b200.h,
            # ``ready_dev``), so the decode of chunk k overlaps the transfer of chunk k+1.
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            copy = self._copy_stream
            n_chunks = (B + H2D_CHUNK - 1) // H2D_CHUNK
            d_logits = torch.empty(logits.shape, dtype=torch.float32, device=dev)
            d_ready = torch.zeros(1, dtype=torch.int32, device=dev)
            marks = torch.tensor([min(B, (k + 1) * H2D_CHUNK) for k in range(n_chunks)], dtype=torch.int32).pin_memory()
            copy.wait_stream(main)  # buffers allocated / zeroed on the main stream
            with torch.cuda.stream(copy):
                for k in range(n_chunks):
                    a0, b0 = k * H2D_CHUNK, min(B, (k + 1) * H2D_CHUNK)
                    d_logits[a0:b0].copy_(logits[a0:b0], non_blocking=True)
                    d_ready.copy_(marks[k:k + 1], non_blocking=True)
            d_logits.record_stream(copy)
            d_ready.record_stream(copy)
            # chunk-major work order, longest first inside a chunk
            chunk_id = torch.arange(B, device=dev, dtype=torch.int64) // H2D_CHUNK
            d_order = torch.argsort(chunk_id * (int(logits.shape[1]) + 1) - d_len.to(torch.int64)).to(torch.int32)
            self._keepalive = (logits, marks)  # host buffers stay alive until the copies ran
            d_n, d_logit, d_comb, d_tok, d_lens, d_status, *wf = self.decode_launch(
                d_logits, d_len, d_order, beam_width, beam_prune_logp, token_min_logp, n_best, input_mode, d_stats,
                ready=(d_ready, H2D_CHUNK), word_frames=word_frames, prune_history=prune_history)
        else:
            d_logits = logits.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            d_order = torch.argsort(d_len, descending=True).to(torch.int32)
            d_n, d_logit, d_comb, d_tok, d_lens, d_status, *wf = self.decode_launch(
                d_logits, d_len, d_order, beam_width, beam_prune_logp, token_min_logp, n_best, input_mode, d_stats,
                word_frames=word_frames, prune_history=prune_history)
        if not to_host:
            return (d_n, d_logit, d_comb, d_tok, d_lens, d_status, d_stats, *wf)
        return self._to_host((d_n, d_logit, d_comb, d_tok, d_lens, d_status, *wf), d_stats)

    @staticmethod
    def _to_host(outs, d_stats=None) -> DecodedBatch:
        d_n, d_logit, d_comb, d_tok, d_lens, d_status, *wf = outs
        out = DecodedBatch(d_n.cpu().numpy(), d_logit.cpu().numpy(), d_comb.cpu().numpy(), d_tok.cpu().numpy(),
                           d_lens.cpu().numpy(), d_status.cpu().numpy(),
                           d_stats.cpu().numpy() if d_stats is not None else None,
                           *(t.cpu().numpy() for t in wf))
        if out.status.any():
            bad = np.nonzero(out.status)[0]
            raise _lib.CoralError(int(out.status[bad[0]]), f"decoder arena capacity exceeded for utterances {bad[:8].tolist()}")
        return out

    def decode_launch(self, d_logits, d_len, d_order, beam_width: int = DEFAULT_BEAM_WIDTH,
                      beam_prune_logp: float = DEFAULT_PRUNE_LOGP, token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP,
                      n_best: int = 1, input_mode: int = 0, d_stats=None, events=None, ready=None,
                      word_frames: bool = False, prune_history: bool = False):
        """Queue one batched decode on device-resident inputs (asynchronous). ``events`` =
        (start, end) ``torch.cuda.Event`` recorded around the library call on the launching stream.
        ``ready`` = (int32 device counter, chunk) for logits still arriving on another stream."""
        B, T_max, V = d_logits.shape
        return self._launch_raw(d_logits.data_ptr(), B, int(T_max), int(V), d_logits.device, d_len, d_order,
                                beam_width, beam_prune_logp, token_min_logp, n_best, input_mode, d_stats, events,
                                ready, word_frames, prune_history)

    def _launch_raw(self, logits_ptr: int, B: int, T_max: int, V: int, dev, d_len, d_order,
                    beam_width: int = DEFAULT_BEAM_WIDTH, beam_prune_logp: float = DEFAULT_PRUNE_LOGP,
                    token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP, n_best: int = 1, input_mode: int = 0,
                    d_stats=None, events=None, ready=None, word_frames: bool = False, prune_history: bool = False,
                    d_frame_off=None):
        """``logits_ptr``: device memory, or pinned host memory (read in place). ``d_frame_off``:
        int64 [B] first frame of each utterance in a packed ``[sum T, V]`` buffer (ragged input)."""
        torch = _torch()
        h = self._handle()
        Tm = max(int(T_max), 1)
        d_n = torch.empty(B, dtype=torch.int32, device=dev)
        d_logit = torch.empty((B, n_best), dtype=torch.float64, device=dev)
        d_comb = torch.empty((B, n_best), dtype=torch.float64, device=dev)
        d_tok = torch.zeros((B, n_best, Tm), dtype=torch.uint8, device=dev)
        d_lens = torch.zeros((B, n_best), dtype=torch.int32, device=dev)
        d_status = torch.empty(B, dtype=torch.int32, device=dev)
        d_wf = d_wn = None
        max_words = (Tm + 1) // 2 + 1  # a word needs a letter frame and (all but the last) a space frame
        if word_frames:
            d_wf = torch.empty((B, n_best, max_words, 2), dtype=torch.int32, device=dev)
            d_wn = torch.zeros((B, n_best), dtype=torch.int32, device=dev)
        if events is not None:
            events[0].record()
        _lib.check(_lib.load().coral_ctc_beam_decode(
            h, logits_ptr, d_len.data_ptr(), d_order.data_ptr() if d_order is not None else None,
            d_frame_off.data_ptr() if d_frame_off is not None else None, B,
            int(T_max), V, int(beam_width), float(beam_prune_logp), float(token_min_logp), int(bool(prune_history)),
            int(input_mode),
            int(n_best), d_n.data_ptr(), d_logit.data_ptr(), d_comb.data_ptr(), d_tok.data_ptr(), d_lens.data_ptr(),
            d_status.data_ptr(), d_stats.data_ptr() if d_stats is not None else None,
            _lib.ptr(ready[0]) if ready is not None else None, int(ready[1]) if ready is not None else 0,
            d_wf.data_ptr() if word_frames else None, d_wn.data_ptr() if word_frames else None, max_words,
            _lib.stream_ptr(dev)))
        if events is not None:
            events[1].record()
        if word_frames:
            return d_n, d_logit, d_comb, d_tok, d_lens, d_status, d_wf, d_wn
        return d_n, d_logit, d_comb, d_tok, d_lens, d_status

    # ------------------------------------------------------------ host rows -> ragged pinned buffer
    def _staging(self, n_floats: int):
        """Pinned staging buffer of the next slot (two slots alternate, so that one batch can be packed
        while the launch of the previous one still reads its own; grown geometrically, reused call
        after call). The launch that last read this slot is waited for before it is overwritten.
        Returns (buffer, slot)."""
        torch = _torch()
        if getattr(self, "_stage", None) is None:
            self._stage = [{"buf": None, "event": None, "ready": None} for _ in range(2)]
            self._stage_next = 0
        slot = self._stage_next
        self._stage_next = 1 - slot
        st = self._stage[slot]
        if st["event"] is not None:
            st["event"].synchronize()
            st["event"] = None
        if st["buf"] is None or st["buf"].numel() < n_floats:
            st["buf"] = torch.empty(max(n_floats, 1 << 16) * 5 // 4, dtype=torch.float32, pin_memory=True)
        return st["buf"], slot

    def _decode_rows(self, ptrs, lens_np, T_pitch, dev, beam_width, beam_prune_logp, token_min_logp, n_best,
                     input_mode, d_stats, word_frames, prune_history):
        """Decode utterances given as host row pointers (``ptrs`` int64 [B] addresses of ``[T_i, V]``
        float32 C-contiguous blocks, ``lens_np`` int64 [B] frames). The rows are packed into the
        pinned ragged staging buffer by host threads and read from there by the kernel (zero-copy).
        Large batches are packed ``H2D_CHUNK`` utterances at a time while the kernel already runs:
        a pinned host counter tells the thread groups which chunks have landed."""
        torch = _torch()
        lib = _lib.load()
        B = len(lens_np)
        V = len(self._idx2vocab)
        row_bytes = V * 4
        off = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(lens_np, out=off[1:])
        stage, slot = self._staging(int(off[-1]) * V)
        dst_off = off[:-1] * row_bytes
        nbytes = lens_np * row_bytes
        ptrs = np.ascontiguousarray(ptrs, dtype=np.int64)
        d_len = torch.from_numpy(lens_np.astype(np.int32)).to(dev, non_blocking=True)
        d_foff = torch.from_numpy(off[:-1].copy()).to(dev, non_blocking=True)
        nthr = _host_threads()
        n_chunks = (B + H2D_CHUNK - 1) // H2D_CHUNK
        Tm = max(int(T_pitch), int(lens_np.max()) if B else 0, 1)
        if n_chunks < 2:
            _lib.check(lib.coral_host_pack_rows(ptrs.ctypes.data, nbytes.ctypes.data, dst_off.ctypes.data, B,
                                                stage.data_ptr(), nthr))
            d_order = torch.argsort(d_len, descending=True).to(torch.int32)
            out = self._launch_raw(stage.data_ptr(), B, Tm, V, dev, d_len, d_order, beam_width, beam_prune_logp,
                                   token_min_logp, n_best, input_mode, d_stats, word_frames=word_frames,
                                   prune_history=prune_history, d_frame_off=d_foff)
        else:
            if self._stage[slot]["ready"] is None:
                self._stage[slot]["ready"] = torch.zeros(16, dtype=torch.int32).pin_memory()
            ready = self._stage[slot]["ready"]
            ready_np = ready.numpy()
            ready_np[0] = 0
            chunk_id = np.arange(B, dtype=np.int64) // H2D_CHUNK
            order = np.argsort(chunk_id * (Tm + 1) - lens_np, kind="stable").astype(np.int32)
            d_order = torch.from_numpy(order).to(dev, non_blocking=True)
            # chunk 0 is packed before the launch: nobody starts by waiting
            a0, b0 = 0, min(B, H2D_CHUNK)
            _lib.check(lib.coral_host_pack_rows(ptrs[a0:].ctypes.data, nbytes[a0:].ctypes.data,
                                                dst_off[a0:].ctypes.data, b0 - a0, stage.data_ptr(), nthr))
            ready_np[0] = b0
            out = self._launch_raw(stage.data_ptr(), B, Tm, V, dev, d_len, d_order, beam_width, beam_prune_logp,
                                   token_min_logp, n_best, input_mode, d_stats, ready=(ready, H2D_CHUNK),
                                   word_frames=word_frames, prune_history=prune_history, d_frame_off=d_foff)
            for k in range(1, n_chunks):
                a0, b0 = k * H2D_CHUNK, min(B, (k + 1) * H2D_CHUNK)
                _lib.check(lib.coral_host_pack_rows(ptrs[a0:].ctypes.data, nbytes[a0:].ctypes.data,
                                                    dst_off[a0:].ctypes.data, b0 - a0, stage.data_ptr(), nthr))
                ready_np[0] = b0  # x86 keeps stores in order: the rows are visible before the counter
        self._stage[slot]["event"] = torch.cuda.Event()
        self._stage[slot]["event"].record(torch.cuda.current_stream(dev))
        return out

    def device_text(self, d_tok, d_lens):
        """Winning token rows on the device (``d_tok`` uint8 ``[B, n_best, T]`` + ``d_lens`` int32
        ``[B, n_best]``) -> (code points int32 ``[cap]``, offsets int64 ``[B + 1]``, max_len int32
        scalar), all device tensors: the best beam's text as flat UTF-32 (``coral_decoder_tokens_to_text``)."""
        torch = _torch()
        B, nb, T = d_tok.shape
        dev = d_tok.device
        maxcp = max([len(c) for c in self._alphabet.labels] + [1])
        cap = max(1, B * T * maxcp)
        d_cps = torch.empty(cap, dtype=torch.int32, device=dev)
        d_off = torch.empty(B + 1, dtype=torch.int64, device=dev)
        d_work = torch.empty(max(B, 1), dtype=torch.int64, device=dev)
        d_max = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().coral_decoder_tokens_to_text(
            self._handle(), d_tok.data_ptr(), nb * T, d_lens.data_ptr(), nb, B, d_cps.data_ptr(), cap,
            d_off.data_ptr(), d_work.data_ptr(), d_max.data_ptr(), _lib.stream_ptr(dev)))
        return d_cps, d_off, d_max

    def device_tokens_to_text(self, d_tok, d_lens) -> list[str]:
        """``[B, T]`` uint8 token rows + ``[B]`` lengths on the device -> Python strings (the text is
        assembled on the GPU; the host decodes one flat UTF-32 buffer and slices it)."""
        return self._texts_from_device(*self.device_text(d_tok[:, None, :], d_lens[:, None]))

    def _texts_from_device(self, d_cps, d_off, d_max, h_off=None, h_max=None) -> "DecodedTexts":
        """Device text -> list of str. ``h_off`` / ``h_max``: the offsets and the longest length if they
        have already been copied to the host (``_batch_launch`` queues those copies behind the kernels)."""
        B = d_off.numel() - 1
        off = d_off.cpu() if h_off is None else h_off     # sync #1: offsets (and thereby the total)
        o = off.tolist()
        total = o[-1]
        cps = d_cps[:total].cpu().numpy()                 # sync #2: exactly the code points
        # one C loop over the flat buffer builds the list of str (coral_py_string_list); every label
        # below U+0100 (CoRal's alphabet) goes through the one-byte kind, which CPython copies fastest
        off_np = off.numpy()
        if self._latin1:
            flat = cps.astype(np.uint8)
            strings = _lib.load().coral_py_string_list(flat.ctypes.data, 1, off_np.ctypes.data, B)
        else:
            flat = np.ascontiguousarray(cps.view(np.uint32))
            strings = _lib.load().coral_py_string_list(flat.ctypes.data, 4, off_np.ctypes.data, B)
        out = DecodedTexts(strings)
        max_len = (int(d_max.item()) if h_max is None else int(h_max)) if B else 0
        out._coral_dev = (d_cps[:total], d_off, max_len, ("decoded", next(DecodedTexts._tokens)))
        return out

    def tokens_to_text(self, tokens: np.ndarray, lens: np.ndarray) -> list[str]:
        """Alphabet indices -> strings for ``[N, T]`` token rows with ``[N]`` lengths (host arrays)."""
        N, T = tokens.shape
        lens = lens.astype(np.int64)
        mask = np.arange(T)[None, :] < lens[:, None]
        flat = tokens[mask]
        off = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        if self._single_cp or not np.isin(flat, np.nonzero(self._cp_table == 0)[0]).any():
            text = self._cp_table[flat].tobytes().decode("utf-32-le")
            o = off.tolist()
            return [text[a:b] for a, b in zip(o[:-1], o[1:])]
        labels = self._alphabet.labels
        return ["".join(labels[t] for t in flat[off[i] : off[i + 1]]) for i in range(N)]

    def _rows(self, logits_list):
        """list of ``[T_i, V]`` arrays -> (row pointers int64 [B], frames int64 [B], keep-alive list).
        One C loop over the list reads every float32 C-contiguous array through the buffer protocol
        (``coral_py_logits_rows``); only what it does not recognise goes through numpy here."""
        V = len(self._idx2vocab)
        keep = logits_list if isinstance(logits_list, list) else list(logits_list)
        B = len(keep)
        ptrs = np.empty(B, dtype=np.int64)
        lens = np.empty(B, dtype=np.int64)
        other = np.empty(B, dtype=np.uint8)
        if _lib.load().coral_py_logits_rows(keep, V, ptrs.ctypes.data, lens.ctypes.data, other.ctypes.data) != B:
            raise RuntimeError("coral_py_logits_rows failed")
        rest = np.nonzero(other)[0]
        if len(rest):
            keep = list(keep)
            for i in rest.tolist():
                lg = keep[i]
                if not isinstance(lg, np.ndarray):
                    lg = np.asarray(lg)
                self._check_logits_dimension(lg)
                if lg.dtype != np.float32 or not lg.flags.c_contiguous:
                    lg = np.ascontiguousarray(lg, dtype=np.float32)
                keep[i] = lg
                ptrs[i] = lg.ctypes.data
                lens[i] = lg.shape[0]
        return ptrs, lens, keep

    # ---------------------------------------------------- pyctcdecode's public API
    def decode_beams(self, logits, beam_width: int = DEFAULT_BEAM_WIDTH, beam_prune_logp: float = DEFAULT_PRUNE_LOGP,
                     token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP, prune_history: bool = DEFAULT_PRUNE_BEAMS,
                     hotwords=None, hotword_weight: float = DEFAULT_HOTWORD_WEIGHT, lm_start_state=None):
        """All final beams of one utterance: ``[(text, lm_state, [(word, (start, end))],
        logit_score, lm_score)]``, best first."""
        return self._decode_beams_many([logits], beam_width, beam_prune_logp, token_min_logp, prune_history,
                                       hotwords, lm_start_state, n_best=beam_width, mp_safe=False)[0]

    def _decode_beams_many(self, logits_list, beam_width, beam_prune_logp, token_min_logp, prune_history,
                           hotwords, lm_start_state, n_best, mp_safe):
        if hotwords:
            raise NotImplementedError("hotwords are not implemented (never passed by CoRal; SURVEY.md 8 A9)")
        if lm_start_state is not None:
            raise NotImplementedError("lm_start_state (stateful decoding) is not implemented")
        logits_list = list(logits_list)
        if not logits_list:
            return []
        if not 1 <= beam_width <= MAX_BEAM_WIDTH:
            raise ValueError(f"beam_width must be in [1, {MAX_BEAM_WIDTH}]")
        torch = _torch()
        self._handle()
        dev = torch.device("cuda", self._device)
        ptrs, lens, keep = self._rows(logits_list)
        n_best = max(1, min(int(n_best), int(beam_width)))
        T_max = max(int(lens.max()), 1)
        # bound the word-frame output buffer (B x n_best x max_words x 8 bytes) per launch
        per_utt = n_best * ((T_max + 1) // 2 + 1) * 8 + n_best * T_max
        step = max(1, min(len(logits_list), (1 << 30) // max(per_utt, 1)))
        res = []
        for a0 in range(0, len(logits_list), step):
            outs = self._decode_rows(ptrs[a0:a0 + step], lens[a0:a0 + step], 0, dev, beam_width, beam_prune_logp,
                                     token_min_logp, n_best, 0, None, True, prune_history)
            out = self._to_host(outs)
            B, nb = out.lens.shape
            texts = self.tokens_to_text(out.tokens.reshape(B * nb, -1), out.lens.reshape(-1))
            for u in range(B):
                beams = []
                for r in range(min(int(out.n_beams[u]), nb)):
                    text = texts[u * nb + r]
                    wf = out.word_frames[u, r, : out.word_counts[u, r]]
                    # pyctcdecode: list(zip(text.split(), text_frames))
                    frames = list(zip(text.split(), ((int(x), int(y)) for x, y in wf)))
                    ls, cs = float(out.logit_score[u, r]), float(out.lm_score[u, r])
                    beams.append((text, frames, ls, cs) if mp_safe else (text, None, frames, ls, cs))
                res.append(beams)
        del keep
        return res

    def decode(self, logits, beam_width: int = DEFAULT_BEAM_WIDTH, beam_prune_logp: float = DEFAULT_PRUNE_LOGP,
               token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP, hotwords=None,
               hotword_weight: float = DEFAULT_HOTWORD_WEIGHT, lm_start_state=None) -> str:
        if lm_start_state is not None:
            raise NotImplementedError("lm_start_state (stateful decoding) is not implemented")
        return self.decode_batch(None, [logits], beam_width, beam_prune_logp, token_min_logp, hotwords,
                                 hotword_weight)[0]

    def decode_beams_batch(self, pool, logits_list, beam_width: int = DEFAULT_BEAM_WIDTH,
                           beam_prune_logp: float = DEFAULT_PRUNE_LOGP, token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP,
                           prune_history: bool = False, hotwords=None,
                           hotword_weight: float = DEFAULT_HOTWORD_WEIGHT, n_best: int | None = None):
        """MP-safe beams (4-tuples: text, frames, logit_score, lm_score) for every utterance.
        ``pool`` is ignored. ``n_best`` (extension) caps the beams returned per utterance."""
        return self._decode_beams_many(list(logits_list), beam_width, beam_prune_logp, token_min_logp,
                                       prune_history, hotwords, None,
                                       n_best=beam_width if n_best is None else n_best, mp_safe=True)

    def decode_batch(self, pool, logits_list, beam_width: int = DEFAULT_BEAM_WIDTH,
                     beam_prune_logp: float = DEFAULT_PRUNE_LOGP, token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP,
                     hotwords=None, hotword_weight: float = DEFAULT_HOTWORD_WEIGHT, lengths=None) -> list[str]:
        """Best transcript of every utterance. ``logits_list`` is a list of ``[T_i, V]``
        arrays, or (extension) a padded ``[B, T_max, V]`` array/tensor with ``lengths``."""
        if hotwords:
            raise NotImplementedError("hotwords are not implemented (never passed by CoRal; SURVEY.md 8 A9)")
        launched = self._batch_launch(logits_list, lengths, beam_width, beam_prune_logp, token_min_logp)
        return [] if launched is None else self._batch_finish(launched)

    def _batch_launch(self, logits_list, lengths, beam_width, beam_prune_logp, token_min_logp):
        """Queue one batch on the current stream: decode + transcripts as UTF-32 on the device. Returns the
        device tensors (and what must stay alive until the launch has run), or None for an empty list."""
        torch = _torch()
        if lengths is None:
            logits_list = list(logits_list)
            if not logits_list:
                return None
            if not 1 <= beam_width <= MAX_BEAM_WIDTH:
                raise ValueError(f"beam_width must be in [1, {MAX_BEAM_WIDTH}]")
            self._handle()
            dev = torch.device("cuda", self._device)
            ptrs, lens, keep = self._rows(logits_list)
            outs = self._decode_rows(ptrs, lens, 0, dev, beam_width, beam_prune_logp, token_min_logp, 1, 0, None,
                                     False, False)
            d_tok, d_lens, d_status = outs[3], outs[4], outs[5]
        else:
            d_n, d_logit, d_comb, d_tok, d_lens, d_status, _ = self.decode_padded(
                logits_list, lengths, beam_width, beam_prune_logp, token_min_logp, n_best=1, to_host=False)
            keep = (getattr(self, "_keepalive", None), logits_list, lengths)  # inputs outlive the launch
        d_cps, d_off, d_max = self.device_text(d_tok, d_lens)
        # the small read-backs (capacity flag, longest transcript, offsets) are queued right behind the
        # kernels, on the same stream, into pinned memory: collecting the batch later needs no kernel
        # of its own -- which, behind a persistent decode of the NEXT batch, would wait for a free SM
        B = d_off.numel() - 1
        bad = d_status.abs().max().to(torch.int32).reshape(1) if d_status.numel() else \
            torch.zeros(1, dtype=torch.int32, device=d_off.device)
        h_small = torch.empty(2, dtype=torch.int32, pin_memory=True)
        h_small.copy_(torch.cat([bad, d_max.reshape(1).to(torch.int32)]), non_blocking=True)
        h_off = torch.empty(B + 1, dtype=torch.int64, pin_memory=True)
        h_off.copy_(d_off, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(d_off.device))
        return (d_cps, d_off, d_max, d_status, keep, h_small, h_off, done)

    def _batch_finish(self, launched) -> "DecodedTexts":
        d_cps, d_off, d_max, d_status, _keep, h_small, h_off, done = launched
        done.synchronize()
        if int(h_small[0]):  # some utterance ran out of arena capacity
            status = d_status.cpu().numpy()
            bad = np.nonzero(status)[0]
            raise _lib.CoralError(int(status[bad[0]]), f"decoder arena capacity exceeded for utterances {bad[:8].tolist()}")
        return self._texts_from_device(d_cps, d_off, d_max, h_off, int(h_small[1]))

    def decode_batches(self, batches, beam_width: int = DEFAULT_BEAM_WIDTH,
                       beam_prune_logp: float = DEFAULT_PRUNE_LOGP, token_min_logp: float = DEFAULT_MIN_TOKEN_LOGP,
                       prefetch: int = 2):
        """(extension) ``decode_batch`` over an iterable of batches, ``prefetch`` batches ahead: yields
        the transcripts of batch k while batch k + 1 is decoding and batch k + 2 is queued behind it, so
        the caller's host work on batch k (strings, ``cer`` / ``wer``, bookkeeping) overlaps the next
        decode and the GPU goes from one decode straight into the next (the small kernels the caller
        queues for batch k run in the tail of decode k + 1) -- the loop
        ``for batch in dataloader: decode; score`` of R:src/coral/evaluate.py / validation.py.

        Every item of ``batches`` is either a list of ``[T_i, V]`` arrays or a ``(padded logits,
        lengths)`` pair, as ``decode_batch`` takes them. The decodes are queued on a stream this
        decoder owns; the stream that is current when a batch is yielded waits for exactly that batch,
        so work the caller queues (the metric kernels) is not held up behind the next decode.
        Results and their order are those of calling ``decode_batch`` batch by batch."""
        import collections

        torch = _torch()
        self._handle()
        dev = torch.device("cuda", self._device)
        if getattr(self, "_pipe_stream", None) is None:
            self._pipe_stream = torch.cuda.Stream(device=dev)
        pipe = self._pipe_stream
        it = iter(batches)
        pending = collections.deque()

        def launch():
            try:
                batch = next(it)
            except StopIteration:
                return False
            logits, lengths = batch if isinstance(batch, tuple) and len(batch) == 2 else (batch, None)
            pipe.wait_stream(torch.cuda.current_stream(dev))  # inputs produced on the caller's stream
            with torch.cuda.stream(pipe):
                launched = self._batch_launch(logits, lengths, beam_width, beam_prune_logp, token_min_logp)
            pending.append(launched)
            return True

        more = True
        depth = 1 + max(0, int(prefetch))
        try:
            while True:
                while more and len(pending) < depth:
                    more = launch()
                if not pending:
                    break
                launched = pending.popleft()
                if launched is None:
                    yield []
                    continue
                cur = torch.cuda.current_stream(dev)
                cur.wait_event(launched[-1])
                for t in launched[:4]:
                    t.record_stream(cur)  # allocated on the decoder's stream, consumed on the caller's
                yield self._batch_finish(launched)
        finally:
            # a caller that stops early (break, exception) leaves decodes in flight that read the
            # caller's buffers in place: wait for them before those buffers can go away
            if any(p is not None for p in pending):
                pipe.synchronize()
            pending.clear()

    # ------------------------------------------------------------- serialisation
    def save_to_dir(self, filepath: str) -> None:
        os.makedirs(filepath, exist_ok=True)
        with open(os.path.join(filepath, self._ALPHABET_SERIALIZED_FILENAME), "w") as fi:
            fi.write(self._alphabet.dumps())
        lm = self._language_model
        if lm is None:
            logger.info("decoder has no language model.")
        else:
            lm.save_to_dir(os.path.join(filepath, self._LANGUAGE_MODEL_SERIALIZED_DIRECTORY))

    @staticmethod
    def parse_directory_contents(filepath: str) -> dict:
        contents = os.listdir(filepath)
        if BeamSearchDecoderCTC._ALPHABET_SERIALIZED_FILENAME not in contents:
            raise ValueError(f"Could not find alphabet file {BeamSearchDecoderCTC._ALPHABET_SERIALIZED_FILENAME}")
        lm_dir = os.path.join(filepath, BeamSearchDecoderCTC._LANGUAGE_MODEL_SERIALIZED_DIRECTORY)
        return {
            "alphabet": os.path.join(filepath, BeamSearchDecoderCTC._ALPHABET_SERIALIZED_FILENAME),
            "language_model": lm_dir if os.path.isdir(lm_dir) else None,
        }

    @classmethod
    def load_from_dir(cls, filepath: str, unigram_encoding: str | None = None) -> "BeamSearchDecoderCTC":
        filenames = cls.parse_directory_contents(filepath)
        with open(filenames["alphabet"]) as fi:
            alphabet = Alphabet.loads(fi.read())
        if filenames["language_model"] is None:
            language_model = None
        else:
            language_model = LanguageModel.load_from_dir(filenames["language_model"], unigram_encoding)
        return cls(alphabet, language_model=language_model)

    @classmethod
    def load_from_hf_hub(cls, model_id: str, cache_dir=None, **kwargs) -> "BeamSearchDecoderCTC":
        from huggingface_hub import snapshot_download

        cached = snapshot_download(model_id, cache_dir=cache_dir, **kwargs)
        return cls.load_from_dir(cached)


def build_ctcdecoder(labels, kenlm_model_path: str | None = None, unigrams=None, alpha: float = DEFAULT_ALPHA,
                     beta: float = DEFAULT_BETA, unk_score_offset: float = DEFAULT_UNK_LOGP_OFFSET,
                     lm_score_boundary: bool = DEFAULT_SCORE_LM_BOUNDARY) -> BeamSearchDecoderCTC:
    """UP:pyctcdecode ``decoder.build_ctcdecoder`` as CoRal calls it at R:src/coral/ngram.py:341-343."""
    alphabet = Alphabet.build_alphabet(list(labels))
    if kenlm_model_path is None:
        return BeamSearchDecoderCTC(alphabet, None)
    kenlm_model = KenlmModel(kenlm_model_path)
    if str(kenlm_model_path).endswith(".arpa") and unigrams is None:
        unigrams = load_unigram_set_from_arpa(kenlm_model_path)
    language_model = LanguageModel(kenlm_model, unigrams, alpha=alpha, beta=beta,
                                   unk_score_offset=unk_score_offset, score_boundary=lm_score_boundary)
    return BeamSearchDecoderCTC(alphabet, language_model)
