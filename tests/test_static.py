"""Static checks of the GPU-only Python paths (they never run in the CPU suite): every name that
is loaded is defined somewhere in its module."""

from __future__ import annotations

import ast
import builtins
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _undefined(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    defined = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            defined.add(n.name)
        elif isinstance(n, ast.Import):
            defined.update((a.asname or a.name).split(".")[0] for a in n.names)
        elif isinstance(n, ast.ImportFrom):
            defined.update(a.asname or a.name for a in n.names)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            defined.add(n.id)
        elif isinstance(n, ast.arg):
            defined.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            defined.add(n.name)
    return [(n.lineno, n.id) for n in ast.walk(tree)
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in defined]


def test_no_undefined_names_in_product_bench_and_tools():
    files = glob.glob(os.path.join(ROOT, "coral_b200", "**", "*.py"), recursive=True)
    files += [os.path.join(ROOT, f) for f in ("bench.py", "bench_configs.py", "__graft_entry__.py", "synth.py")]
    files += glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py"))
    bad = {os.path.relpath(f, ROOT): u for f in files if (u := _undefined(f))}
    assert not bad, bad
