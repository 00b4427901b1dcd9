"""Generates tests/golden/normaliser_ref.json by running the REFERENCE's own code.

`import coral` fails in the build image (hydra / omegaconf / jiwer are missing), but the text
normaliser is self-contained: this script lifts `NUMERAL_REGEX` + `convert_numeral_to_words` out of
/root/reference/src/coral/utils.py and `DEFAULT_CONVERSION_DICT`, `FILLER_WORDS_PATTERN` +
`process_example` out of /root/reference/src/coral/data.py with `ast` (no other module-level code
of those files runs) and executes them on (a) the known-answer inputs of the reference's own tests
(/root/reference/tests/test_data.py:72-235, tests/test_utils.py:52-126) and (b) seeded random
Danish-looking text with numerals, filler words, compatibility characters, combining marks and
the symbols of the conversion dict, under the option combinations the reference uses. The outputs
are REFERENCE outputs: they pin coral_b200/csrc/normalise.cc (tests/test_normaliser.py).

    python tests/golden/make_normaliser_golden.py        # needs /root/reference; run in the build image
"""

import ast
import json
import logging
import os
import re
import sys
from unicodedata import normalize

import numpy as np

REF = "/root/reference/src/coral"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift(path, names):
    tree = ast.parse(open(path, encoding="utf-8").read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in names for t in node.targets):
            keep.append(node)
    from collections.abc import Callable, Iterable

    ns = {"re": re, "normalize": normalize, "logger": logging.getLogger("ref"), "Iterable": Iterable, "Callable": Callable}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    u = lift(os.path.join(REF, "utils.py"), {"NUMERAL_REGEX", "convert_numeral_to_words"})
    d_ns = lift(os.path.join(REF, "data.py"), {"DEFAULT_CONVERSION_DICT", "FILLER_WORDS_PATTERN", "process_example"})
    d_ns.update(NUMERAL_REGEX=u["NUMERAL_REGEX"], convert_numeral_to_words=u["convert_numeral_to_words"])
    process_example = d_ns["process_example"]
    convert = u["convert_numeral_to_words"]
    DEFAULT = d_ns["DEFAULT_CONVERSION_DICT"]

    def run(text, keep, conv, lower, numerals):
        ex = process_example(example={"text": text}, characters_to_keep=keep, conversion_dict=conv, text_column="text",
                             audio_column=None, lower_case=lower, convert_numerals=numerals, processor=None,
                             normalise_audio=True, augment_audio=False)
        return ex["text"]

    rng = np.random.default_rng(4242)
    coral_chars = "abcdefghijklmnopqrstuvwxyzæøå0123456789éü"   # R:config/evaluation.yaml:14
    words = ["hej", "med", "dig", "Århus", "RØDGRØD", "fløde", "på", "Og", "ehh", "Øhm", "hmm", "mmh", "ehmx", "eh", "hm",
             "f.eks.", "fx", "hhv", "kg", "μg", "aabenraa", "Ærø", "İstanbul", "straße", "ǅ", "ﬁn", "ｆｕｌｌ", "①", "½",
             "x²", "ĳ", "Ǆ", "naïve", "café", "ạ̈", "̣̈o", "가", "각", "—", "–", "…", "“hej”",
             "§", "%", "‰", "+", "-", "50%", "3,5", "1.000", "10.102,92", "1.000000", "05", "007", "12a", "a1", "_5", "2.5",
             "1234567890", "999.999.999", "100", "1000", "21", "0", "0,50", "7,", ",7", "1.00", "12.345.678", "|", "\t", "\n",
             "​", " ", " ", "ÅÄÖ", "þorn", "ð", "ı", "ſ", "K", "Å", "µ", "tést", "(test)", "[x]", "{y}", "!", "?"]
    cases = []
    # (a) the reference's own known answers, recomputed by the reference
    t = "\nThis is a (test) [sentence]́ with \n{aa} and ğ. "
    allc = set(t) | {"å", "g", " "}
    keeps = [sorted(allc), None, sorted(allc - set("()[]{}")), sorted(allc - set("\n\r"))]
    convs = [{}, {"aa": "å", "ğ": "g"}, {"́": " "}]
    for keep in keeps:
        for conv in convs:
            for lower in (True, False):
                cases.append(dict(text=t, keep=keep, conv=conv, lower=lower, numerals=False))
    # (b) seeded random text
    opt = [(list(coral_chars), DEFAULT, True, True), (list(coral_chars), DEFAULT, True, False), (None, DEFAULT, True, True),
           (None, {}, False, True), (None, {}, False, False), (list(coral_chars + "ABCÆ.,"), DEFAULT, False, True),
           (list("abcdefghijklmnopqrstuvwxyzæøå "), {"aa": "å"}, True, True)]
    for k in range(3500):
        n = int(rng.integers(1, 14))
        toks = [words[int(i)] if rng.random() < 0.7 else str(int(rng.integers(0, 10 ** int(rng.integers(1, 10)))))
                for i in rng.integers(0, len(words), size=n)]
        seps = [" ", " ", " ", "  ", "", ", ", ".", "\n", " \n ", "-", " "]
        text = "".join(tok + seps[int(rng.integers(0, len(seps)))] for tok in toks)
        keep, conv, lower, numerals = opt[k % len(opt)]
        cases.append(dict(text=text, keep=keep, conv=conv, lower=lower, numerals=numerals))
    logging.disable(logging.CRITICAL)
    out_cases, options = [], []
    for cse in cases:
        try:
            exp = run(cse["text"], cse["keep"], cse["conv"], cse["lower"], cse["numerals"])
        except Exception as e:  # the reference itself fails on a few inputs (KeyError on exotic digits)
            exp = {"raises": type(e).__name__}
        o = dict(keep=cse["keep"], conv=list(cse["conv"].items()), lower=cse["lower"], numerals=cse["numerals"])
        if o not in options:
            options.append(o)
        out_cases.append([options.index(o), cse["text"], exp])
    # numerals: the reference tests' list + random ones, straight through convert_numeral_to_words
    nums = ["0", "1", "9", "10", "11", "19", "20", "21", "29", "30", "90", "99", "100", "101", "110", "121", "200", "999",
            "1000", "1001", "1010", "1100", "1121", "2000", "10.000", "100.000", "100000", "999.999", "999999", "1.000.000",
            "1.000000", "1.0.00000", "1.000.001", "10.000.000", "100.000.000", "999.999.999", "10,123", "10.102,92",
            "1234567890", "05", "abc", "", "1,0", "1000,01", "2.000.000", "1.001.000", "20.020", "300.003"]
    for _ in range(2500):
        digits = int(rng.integers(1, 11))
        v = str(int(rng.integers(10 ** (digits - 1), 10 ** digits))) if digits > 1 else str(int(rng.integers(0, 10)))
        r = rng.random()
        if r < 0.3 and len(v) > 3:
            v = f"{int(v):,}".replace(",", ".")
        if rng.random() < 0.2:
            v += "," + str(int(rng.integers(0, 1000)))
        nums.append(v)
    logging.disable(logging.CRITICAL)
    num_cases = [[v, convert(numeral=v)] for v in nums]
    import gzip

    with gzip.open(os.path.join(HERE, "normaliser_ref.json.gz"), "wt", encoding="utf-8", compresslevel=9) as f:
        json.dump({"generator": "tests/golden/make_normaliser_golden.py executing /root/reference/src/coral/{data,utils}.py",
                   "python": sys.version.split()[0], "options": options, "process_example": out_cases,
                   "convert_numeral_to_words": num_cases}, f, ensure_ascii=True, separators=(",", ":"))
    print(len(out_cases), "process_example cases under", len(options), "option sets,", len(num_cases), "numerals,",
          sum(isinstance(c[2], dict) for c in out_cases), "where the reference raises")


if __name__ == "__main__":
    main()
