"""Strings <-> UTF-32 code-point buffers for the C ABI (strings cross it as cps + offsets)."""

from __future__ import annotations

import numpy as np


def encode_utf32(strings) -> tuple[np.ndarray, np.ndarray]:
    """``(cps uint32 [sum len], offsets int64 [n + 1])`` for a sequence of ``str``."""
    strings = strings if isinstance(strings, (list, tuple)) else list(strings)
    n = len(strings)
    offsets = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter(map(len, strings), dtype=np.int64, count=n), out=offsets[1:])
    joined = "".join(strings)
    cps = np.frombuffer(bytearray(joined.encode("utf-32-le", "surrogatepass")), dtype=np.uint32)
    if cps.size == 0:
        cps = np.zeros(1, dtype=np.uint32)  # never hand out a NULL pointer
    return cps, offsets


def decode_utf32(cps: np.ndarray, offsets: np.ndarray) -> list[str]:
    text = np.ascontiguousarray(cps, dtype=np.uint32).tobytes().decode("utf-32-le", "surrogatepass")
    o = offsets.tolist()
    return [text[a:b] for a, b in zip(o[:-1], o[1:])]
