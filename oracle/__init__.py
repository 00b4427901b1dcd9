"""CPU oracle for the CTC-decode + WER/CER hot path of alexandrainst/coral.

THIS PACKAGE IS TEST INFRASTRUCTURE. It is a plain-Python / numpy restatement of
the reference's algorithm for the path and exists only to *check* the CUDA path.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it. Nothing under
``coral_b200/`` imports it, and the product path raises when the CUDA library is
missing instead of falling back to anything in here.

PARITY STATUS -- read before trusting it:

* The arithmetic of the reference's hot path lives in third-party packages that
  are neither vendored under ``/root/reference`` nor installed in this image
  (no network): ``pyctcdecode 0.5.0`` (R:uv.lock:2357-2358), ``kenlm 0.2.0``
  from github master.zip (R:uv.lock:1275-1278), ``jiwer 4.0.0``
  (R:uv.lock:1204-1205) -> ``rapidfuzz 3.14.3`` (R:uv.lock:2611-2612) and
  ``pygtrie 2.5.0`` (R:uv.lock:2459-2460). The reference's own tests hold no
  golden vector, known-answer test or fixture for this path.
  The beam-search, n-gram and edit-distance modules are therefore restatements
  of those packages' published algorithms, anchored on the reference's call
  sites (R:src/coral/metrics.py:26-33, :54-61; R:src/coral/ngram.py:336-343;
  HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:359-443):
  **parity unpinned** for ``oracle.beam``, ``oracle.arpa``, ``oracle.lm`` and
  ``oracle.edit``.
* The greedy path IS pinned: ``oracle.greedy`` is checked in ``tests/`` against
  the real ``transformers 5.5.0`` ``Wav2Vec2CTCTokenizer`` (the version the
  reference pins, R:uv.lock:3290-3291), and ``tests/golden/`` holds vectors
  generated from it (script committed beside them).
* ``oracle.selfcheck`` re-runs the restatements against the real packages on any
  machine where ``import pyctcdecode, kenlm, jiwer`` succeeds.

Citation prefixes: ``R:`` = /root/reference, ``HF:`` = the installed
transformers 5.5.0 tree, ``UP:`` = upstream third-party source (not on disk).
"""
