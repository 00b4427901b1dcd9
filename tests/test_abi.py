"""The C-ABI library loads on a CPU-only machine and exports every symbol that
include/coral_b200.h declares (no compute calls here)."""

from __future__ import annotations

import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "coral_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(coral_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("coral_last_error", "coral_lm_load_arpa", "coral_lm_free", "coral_decoder_create",
                 "coral_decoder_set_params", "coral_ctc_beam_decode", "coral_ctc_greedy", "coral_ctc_collapse",
                 "coral_edit_counts", "coral_edit_counts_spans"):
        assert must in names


def test_library_builds_loads_and_exports_all_declared_symbols():
    from coral_b200 import _build, _lib

    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), f"{name} is declared in include/coral_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(_declared())
    loaded = _lib.load()
    assert loaded.coral_abi_version() == 1
    assert loaded.coral_last_error() is not None


def test_ctypes_signatures_have_the_declared_arity():
    """The ctypes binding (coral_b200/_lib.py) and the header agree on the number of parameters of
    every entry point -- the ABI grew several times and a stale binding corrupts the stack silently."""
    from coral_b200 import _lib

    text = open(os.path.join(ROOT, "include", "coral_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, params in re.findall(r"\b(coral_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(_lib.SIGNATURES[name][1]) == n, f"{name}: header has {n} parameters, ctypes binding {len(_lib.SIGNATURES[name][1])}"


def test_argument_errors_map_to_the_reference_exceptions():
    """Status codes -> the exceptions the reference's callers see; pure host paths only."""
    import pytest

    from coral_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    with pytest.raises(OSError):
        _lib.check(lib.coral_lm_load_arpa(b"/nonexistent/file.arpa", 0, ctypes.byref(h)))
    with pytest.raises(ValueError):
        _lib.check(lib.coral_decoder_set_params(None, 0.5, 1.5, -10.0, 1))
    with pytest.raises(ValueError):
        _lib.check(lib.coral_edit_counts(None, None, None, None, 4, 1, 10, 0, None, None, None))
    with pytest.raises(ValueError):  # beam_width out of range is rejected before any CUDA call
        _lib.check(lib.coral_ctc_beam_decode(None, None, None, None, None, 1, 1, 46, 100, -10.0, -5.0, 0, 0, 1,
                                             None, None, None, None, None, None, None, None, 0, None, None, 0, None))


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    """The package must not reference oracle/ or the test-only host simulation."""
    pkg = os.path.join(ROOT, "coral_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "hostsim_lib" not in src and "libcoral_hostsim" not in src, f
