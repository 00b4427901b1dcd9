"""On-device hand-off from the acoustic model to the decoder.

The stock ``AutomaticSpeechRecognitionPipeline`` moves every utterance's logits to the host
(``HF:pipelines/base.py`` ``forward`` -> ``_ensure_tensor_on_device(..., "cpu")``), un-batches
them and calls ``decoder.decode_beams`` once per utterance in ``postprocess``
(``HF:pipelines/automatic_speech_recognition.py:612-616``). This subclass decodes the whole
model batch where it is produced: ``_forward`` hands the CUDA logits of the batch to
``BeamSearchDecoderCTC`` in one launch and only the transcripts (and word frames, when
timestamps are requested) travel on. Everything else -- preprocessing, batching, the result
dictionaries -- is the stock pipeline, so the reference's loop
(``R:src/coral/evaluate.py:56-60``, ``R:src/coral/validation.py:114-120``) runs unchanged:

    transcriber = pipeline("automatic-speech-recognition", model=..., device=device,
                           pipeline_class=coral_b200.pipeline.BatchedCTCWithLMPipeline)
    for out in transcriber(KeyDataset(dataset, "audio"), batch_size=16): out["text"]

Like the stock pipeline it decodes every frame the model returns for an item, padding frames
included (the stock ``postprocess`` does not trim them either). Chunked long-form inputs
(``chunk_length_s``) need the logits of several chunks stitched per item and fall back to the
stock host path.
"""

from __future__ import annotations

from transformers.pipelines.automatic_speech_recognition import AutomaticSpeechRecognitionPipeline

from .decoder import BeamSearchDecoderCTC

_TEXT = "coral_b200_text"
_FRAMES = "coral_b200_word_frames"


class BatchedCTCWithLMPipeline(AutomaticSpeechRecognitionPipeline):
    def _sanitize_parameters(self, *args, decoder_kwargs=None, **kwargs):
        pre, fwd, post = super()._sanitize_parameters(*args, decoder_kwargs=decoder_kwargs, **kwargs)
        if decoder_kwargs is not None:
            fwd = dict(fwd, coral_b200_decoder_kwargs=dict(decoder_kwargs))
        return pre, fwd, post

    def _forward(self, model_inputs, return_timestamps=False, coral_b200_decoder_kwargs=None, **generate_kwargs):
        out = super()._forward(model_inputs, return_timestamps=return_timestamps, **generate_kwargs)
        logits = out.get("logits") if isinstance(out, dict) else None
        if (self.type != "ctc_with_lm" or logits is None or out.get("stride") is not None
                or not isinstance(self.decoder, BeamSearchDecoderCTC) or logits.device.type != "cuda"):
            return out
        import torch

        kw = dict(coral_b200_decoder_kwargs or {})
        if kw.pop("hotwords", None):
            raise NotImplementedError("hotwords are not implemented (never passed by CoRal; SURVEY.md 8 A9)")
        kw.pop("hotword_weight", None)
        B, T, _ = logits.shape
        lengths = torch.full((B,), T, dtype=torch.int32, device=logits.device)
        res = self.decoder.decode_padded(logits.to(torch.float32), lengths, n_best=1,
                                         word_frames=bool(return_timestamps), **kw)
        texts = self.decoder.tokens_to_text(res.tokens[:, 0, :], res.lens[:, 0])
        out[_TEXT] = texts
        if return_timestamps:
            out[_FRAMES] = [
                [(w, (int(a), int(b))) for w, (a, b) in
                 zip(texts[u].split(), res.word_frames[u, 0, : res.word_counts[u, 0]])]
                for u in range(B)
            ]
        # the logits have served their purpose: do not ship them to the host
        out["logits"] = logits.new_zeros((B, 1, 1))
        return out

    def postprocess(self, model_outputs, decoder_kwargs=None, return_timestamps=None, return_language=None):
        if len(model_outputs) != 1 or _TEXT not in model_outputs[0]:
            for o in model_outputs:
                o.pop(_TEXT, None)
                o.pop(_FRAMES, None)
            return super().postprocess(model_outputs, decoder_kwargs=decoder_kwargs,
                                       return_timestamps=return_timestamps, return_language=return_language)
        output = model_outputs[0]
        text = output.pop(_TEXT)
        frames = output.pop(_FRAMES, None)
        if isinstance(text, list):  # a single un-batched item keeps its batch dimension
            text = text[0]
            frames = frames[0] if frames else frames
        optional = {}
        if return_timestamps:
            # same arithmetic as the stock postprocess (HF:...automatic_speech_recognition.py:636-646)
            chunks = []
            for word, (start_offset, end_offset) in frames or []:
                start = start_offset * self._align_to / self.feature_extractor.sampling_rate
                stop = end_offset * self._align_to / self.feature_extractor.sampling_rate
                chunks.append({"text": word, "timestamp": (start, stop)})
            optional["chunks"] = chunks
        for k in ("tokens", "logits", "is_last", "stride", "token_timestamps"):
            output.pop(k, None)
        extra = {k: [v] for k, v in output.items()}
        return {"text": text, **optional, **extra}
