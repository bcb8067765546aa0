"""CPU oracle for the DVC P-frame hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product path (``fastvideocodec_b200``) never imports
this package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the unmodified
reference modules from ``/root/reference`` (with import-time shims only), runs
them on seeded weights and frames and stores the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors (and against the live reference when it is present).
The CompressAI-facing likelihood functions (``eb_*``, ``gaussian_*``) follow
the published CompressAI algorithm; CompressAI is not vendored in the
reference and not installed here, so that one boundary is "parity unpinned".
"""
