"""Text normalisation between decoding and scoring: drop-in for the text part of
``coral.data.process_example`` (R:src/coral/data.py:616-701) and
``coral.utils.convert_numeral_to_words`` (R:src/coral/utils.py:303-472).

``evaluate()`` and ``add_validations()`` push every transcript through ``process_example`` with
``audio_column=None`` (R:src/coral/evaluate.py:61-72, R:src/coral/validation.py:121-132). Once
decoding takes milliseconds, that per-utterance Python (regexes, NFKC, ~40 ``str.replace`` calls)
is the serial tail, so the batch form ``normalise_texts`` runs the same steps in C++ on host
threads (``csrc/normalise.cc``). ``process_example`` keeps the reference's signature for the
text-only case. Nothing here touches the GPU.

Inputs the C++ restatement refuses (a Greek capital sigma under ``lower_case``; a non-ASCII
decimal digit under ``convert_numerals``) raise ``NotImplementedError`` naming the string: they
are never approximated.
"""

from __future__ import annotations

import collections.abc as c
import ctypes as C
import os
import re

import numpy as np

from . import _lib
from .textio import encode_utf32

# R:src/coral/data.py:47-90 -- the characters to convert, applied in this order
DEFAULT_CONVERSION_DICT = {
    "aa": "å", "ğ": "g", "ñ": "n", "ń": "n", "è": "e", "kg": " kilo ", "μg": " mikrogram ",
    "hhv": "henholdsvis", "fx": "for eksempel", "f.eks.": "for eksempel", "-": " minus ", "+": " plus ",
    "μ": " mikro ", "§": " paragraf ", "%": " procent ", "‰": " promille ", "ú": "u", "ş": "s", "ê": "e",
    "ã": "a", "ë": "e", "ć": "c", "ä": "æ", "í": "i", "š": "s", "î": "i", "ě": "e", "ð": "d", "á": "a",
    "ó": "o", "þ": "th", "ı": "i", "ö": "ø", "ç": "c", "ș": "s",
    "\u0301": " ", "\u200b": " ",  # combining acute, zero-width space
}
# R:src/coral/utils.py:31 and R:src/coral/data.py:87-89 (kept for callers that import them)
NUMERAL_REGEX = re.compile(r"\b(0|[1-9]\d{0,2}(?:(?:\.\d{3})*|\d*)(?:,\d+)?)\b")
FILLER_WORDS_PATTERN = re.compile(pattern=r"\b(eh+m*|øh+m*|h+m+|m+h+)\b", flags=re.IGNORECASE)


def _threads() -> int:
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(16, n // max(int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1), 1)))


class TextNormaliser:
    """One configuration of ``process_example`` (characters to keep, conversion dict, flags)."""

    def __init__(self, characters_to_keep: c.Iterable[str] | None, conversion_dict: dict[str, str] | None = None,
                 lower_case: bool = True, convert_numerals: bool = False):
        conversion_dict = DEFAULT_CONVERSION_DICT if conversion_dict is None else conversion_dict
        lib = _lib.load()
        if characters_to_keep is None:
            keep, n_keep = np.zeros(1, np.uint32), -1
        else:
            chars = "".join(ch for ch in characters_to_keep)
            keep = np.array([ord(ch) for ch in chars] or [0], dtype=np.uint32)
            n_keep = len(chars)
        flat = [s for kv in conversion_dict.items() for s in kv]
        cps, off = encode_utf32(flat)
        h = C.c_void_p()
        _lib.check(lib.coral_normaliser_create(keep.ctypes.data, n_keep, cps.ctypes.data, off.ctypes.data,
                                               len(conversion_dict), int(bool(lower_case)),
                                               int(bool(convert_numerals)), C.byref(h)))
        self._h = h

    def __call__(self, texts: list[str], n_threads: int | None = None) -> list[str]:
        texts = list(texts)
        if not texts:
            return []
        lib = _lib.load()
        cps, off = encode_utf32(texts)
        total = C.c_int64()
        _lib.check(lib.coral_normaliser_run(self._h, cps.ctypes.data, off.ctypes.data, len(texts),
                                            n_threads or _threads(), C.byref(total)))
        out_cps = np.empty(max(total.value, 1), dtype=np.uint32)
        out_off = np.empty(len(texts) + 1, dtype=np.int64)
        status = np.zeros(len(texts), dtype=np.int32)
        _lib.check(lib.coral_normaliser_fetch(self._h, out_cps.ctypes.data, out_off.ctypes.data, status.ctypes.data))
        if status.any():
            i = int(np.nonzero(status)[0][0])
            raise NotImplementedError(
                f"text {i} ({texts[i]!r}) holds a Greek capital sigma (lower_case) or a non-ASCII decimal digit "
                "(convert_numerals): not restated by coral_b200's normaliser")
        # one C loop builds the list of str from the flat buffer (as decode_batch does)
        return lib.coral_py_string_list(out_cps.ctypes.data, 4, out_off.ctypes.data, len(texts))

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().coral_normaliser_free(self._h)
                self._h = None
        except Exception:
            pass


def normalise_texts(texts, characters_to_keep: c.Iterable[str] | None, conversion_dict: dict[str, str] | None = None,
                    lower_case: bool = True, convert_numerals: bool = False, n_threads: int | None = None) -> list[str]:
    """Batch form of ``process_example(...)[text_column]`` for ``audio_column=None``."""
    return TextNormaliser(characters_to_keep, conversion_dict, lower_case, convert_numerals)(texts, n_threads)


def convert_numeral_to_words(numeral: str, inside_larger_numeral: bool = False) -> str:
    """R:src/coral/utils.py:303-472. A string that is not (entirely) a numeral comes back unchanged."""
    if inside_larger_numeral:
        raise NotImplementedError("inside_larger_numeral is the reference's recursion flag, not a public input")
    if NUMERAL_REGEX.fullmatch(numeral) is None:
        return numeral
    # the numeral alone through step 1 only: no lowering, no conversion dict, keep everything
    return TextNormaliser(None, {}, lower_case=False, convert_numerals=True)([numeral], n_threads=1)[0]


def process_example(example: dict, characters_to_keep: c.Iterable[str] | None, conversion_dict: dict[str, str],
                    text_column: str, audio_column: str | None, lower_case: bool, convert_numerals: bool,
                    processor=None, normalise_audio: bool = True, augment_audio: bool = False) -> dict:
    """The reference's signature (R:src/coral/data.py:616-627) for the text-only case it is called
    with between decoding and scoring. The audio branch is upstream of the logits and out of scope."""
    if audio_column is not None:
        raise NotImplementedError("only the text-only form (audio_column=None) is provided (SURVEY.md section 8f N3)")
    example[text_column] = normalise_texts([example[text_column]], characters_to_keep, conversion_dict,
                                           lower_case, convert_numerals, n_threads=1)[0]
    return example
