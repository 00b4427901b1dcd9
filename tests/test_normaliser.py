"""The C++ text normaliser (coral_b200/csrc/normalise.cc) against

* the known answers of the reference's OWN tests (/root/reference/tests/test_data.py:72-235 --
  process_example; /root/reference/tests/test_utils.py:52-126 -- convert_numeral_to_words), written
  out below, and
* tests/golden/normaliser_ref.json.gz: ~3500 texts and ~2500 numerals pushed through the reference's
  own functions (lifted from /root/reference/src/coral/{data,utils}.py by
  tests/golden/make_normaliser_golden.py and executed in the build image) -- reference outputs, not
  outputs of a restatement.

Host code only: runs in the CPU suite."""

from __future__ import annotations

import gzip
import json
import os

import pytest

from coral_b200.normalise import (DEFAULT_CONVERSION_DICT, TextNormaliser, convert_numeral_to_words, normalise_texts,
                                  process_example)

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

T = "\nThis is a (test) [sentence]́ with \n{aa} and ğ. "
EMPTY, DIA, WS = {}, {"aa": "å", "ğ": "g"}, {"́": " "}
ALL = set(T) | set(DIA.values()) | set(WS.values())
NO_PAR, NO_NL = ALL - set("()[]{}"), ALL - set("\n\r")


@pytest.mark.parametrize("keep,conv,column,lower,expected", [
    (ALL, EMPTY, "text", True, "this is a (test) [sentence]́ with\n{aa} and ğ."),
    (ALL, EMPTY, "text", False, "This is a (test) [sentence]́ with\n{aa} and ğ."),
    (ALL, EMPTY, "text2", True, "this is a (test) [sentence]́ with\n{aa} and ğ."),
    (None, EMPTY, "text", True, "this is a (test) [sentence]́ with\n{aa} and ğ."),
    (ALL, DIA, "text", True, "this is a (test) [sentence]́ with\n{å} and g."),
    (ALL, WS, "text", True, "this is a (test) [sentence] with\n{aa} and ğ."),
    (NO_PAR, EMPTY, "text", True, "this is a test sentence ́ with\naa and ğ."),
    (NO_PAR, DIA, "text", True, "this is a test sentence ́ with\nå and g."),
    (NO_PAR, WS, "text", True, "this is a test sentence with\naa and ğ."),
    (NO_NL, EMPTY, "text", True, "this is a (test) [sentence]́ with {aa} and ğ."),
    (NO_NL, DIA, "text", True, "this is a (test) [sentence]́ with {å} and g."),
    (NO_NL, WS, "text", True, "this is a (test) [sentence] with {aa} and ğ."),
])
def test_process_example_known_answers_of_the_reference_tests(keep, conv, column, lower, expected):
    out = process_example(example={column: T}, characters_to_keep=keep, conversion_dict=conv, text_column=column,
                          audio_column=None, lower_case=lower, convert_numerals=False, processor=None,
                          normalise_audio=True, augment_audio=False)
    assert out[column] == expected


NUMERALS = [
    ("0", "nul"), ("1", "en"), ("2", "to"), ("3", "tre"), ("4", "fire"), ("5", "fem"), ("6", "seks"), ("7", "syv"),
    ("8", "otte"), ("9", "ni"), ("10", "ti"), ("11", "elleve"), ("12", "tolv"), ("13", "tretten"), ("14", "fjorten"),
    ("15", "femten"), ("16", "seksten"), ("17", "sytten"), ("18", "atten"), ("19", "nitten"), ("20", "tyve"),
    ("21", "enogtyve"), ("22", "toogtyve"), ("23", "treogtyve"), ("24", "fireogtyve"), ("25", "femogtyve"),
    ("26", "seksogtyve"), ("27", "syvogtyve"), ("28", "otteogtyve"), ("29", "niogtyve"), ("30", "tredive"),
    ("40", "fyrre"), ("50", "halvtreds"), ("60", "tres"), ("70", "halvfjerds"), ("80", "firs"), ("90", "halvfems"),
    ("100", "hundrede"), ("101", "et hundrede og en"), ("110", "et hundrede og ti"),
    ("121", "et hundrede og enogtyve"), ("200", "to hundrede"), ("999", "ni hundrede og nioghalvfems"),
    ("1000", "tusind"), ("1001", "et tusind og en"), ("1010", "et tusind og ti"), ("1100", "et tusind et hundrede"),
    ("1121", "et tusind et hundrede og enogtyve"), ("2000", "to tusind"), ("10.000", "ti tusind"),
    ("100.000", "et hundrede tusind"), ("100000", "et hundrede tusind"),
    ("999.999", "ni hundrede og nioghalvfems tusind ni hundrede og nioghalvfems"),
    ("999999", "ni hundrede og nioghalvfems tusind ni hundrede og nioghalvfems"), ("1.000.000", "en million"),
    ("1.000000", "1.000000"), ("1.0.00000", "1.0.00000"), ("1.000.001", "en million og en"),
    ("10.000.000", "ti millioner"), ("100.000.000", "et hundrede millioner"),
    ("999.999.999", "ni hundrede og nioghalvfems millioner ni hundrede og nioghalvfems tusind ni hundrede og nioghalvfems"),
    ("10,123", "ti komma et to tre"), ("10.102,92", "ti tusind et hundrede og to komma ni to"),
]


@pytest.mark.parametrize("numeral,expected", NUMERALS)
def test_convert_numeral_to_words_known_answers_of_the_reference_tests(numeral, expected):
    assert convert_numeral_to_words(numeral=numeral) == expected


def _golden():
    with gzip.open(os.path.join(G, "normaliser_ref.json.gz"), "rt", encoding="utf-8") as f:
        return json.load(f)


def test_against_outputs_of_the_reference_functions():
    data = _golden()
    by_opt = {}
    for o, text, exp in data["process_example"]:
        by_opt.setdefault(o, []).append((text, exp))
    n = 0
    for o, rows in by_opt.items():
        opt = data["options"][o]
        norm = TextNormaliser(opt["keep"], dict(map(tuple, opt["conv"])), opt["lower"], opt["numerals"])
        texts = [t for t, e in rows if not isinstance(e, dict)]
        want = [e for t, e in rows if not isinstance(e, dict)]
        for threads in (1, 4):
            got = norm(texts, n_threads=threads)
            for t, g, w in zip(texts, got, want):
                assert g == w, (opt["lower"], opt["numerals"], t, g, w)
        n += len(texts)
    assert n >= 3000
    for numeral, want in data["convert_numeral_to_words"]:
        assert convert_numeral_to_words(numeral) == want, numeral


def test_refusals_and_defaults():
    with pytest.raises(NotImplementedError):
        normalise_texts(["ΟΔΥΣΣΕΥΣ"], None, {}, lower_case=True)
    with pytest.raises(NotImplementedError):
        normalise_texts(["der er ٣ æbler"], None, {}, lower_case=False, convert_numerals=True)
    with pytest.raises(NotImplementedError):
        process_example({"text": "x", "audio": {}}, None, {}, "text", "audio", True, False)
    assert normalise_texts([], None) == []
    # the evaluation configuration: R:config/evaluation.yaml:14 characters, default conversion dict
    chars = "abcdefghijklmnopqrstuvwxyzæøå0123456789éü"
    out = normalise_texts(["Øhm, det koster 1.250 kr. – ca. 50% af 2.500!", "AARHUS  er f.eks. en by"], chars,
                          DEFAULT_CONVERSION_DICT, lower_case=True, convert_numerals=True)
    assert out == ["det koster et tusind to hundrede og halvtreds kr ca halvtreds procent af to tusind fem hundrede",
                   "århus er for eksempel en by"]
