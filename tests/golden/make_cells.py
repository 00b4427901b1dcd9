"""Generates tests/golden/cells/<cell>.json.gz: the oracle's full beam lists for the parity cells
whose CPython pass is too slow to repeat on every test run (tests/cells.py: CACHED).

    python tests/golden/make_cells.py [cell ...]

Inputs are rebuilt from seeds by tests/cells.py; the file records their SHA-256 fingerprint, so
a test that finds different inputs ignores the file and runs the oracle live instead.
"""

import gzip
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cells  # noqa: E402


def main():
    names = sys.argv[1:] or list(cells.CACHED)
    os.makedirs(cells.CELL_DIR, exist_ok=True)
    for name in names:
        t0 = time.time()
        beams = cells.live_oracle(name)
        data = {"cell": name, "fingerprint": cells.fingerprint(name), "kwargs": cells.CELLS[name]["kw"],
                "workload": cells.CELLS[name]["wl"], "generator": "oracle.beam via tests/golden/make_cells.py",
                "beams": beams}
        with gzip.open(cells.cache_path(name), "wt", encoding="utf-8", compresslevel=9) as f:
            json.dump(data, f, ensure_ascii=False, separators=(",", ":"))
        print(name, "utterances", len(beams), "beams", sum(len(b) for b in beams),
              "%.1f s" % (time.time() - t0), os.path.getsize(cells.cache_path(name)), "bytes", flush=True)


if __name__ == "__main__":
    main()
