"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference algorithm for the VTAMIQ hot path.

Nothing under ``oracle/`` is imported by the ``vtamiq_b200`` package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it, and only as the checker
(or as the timed CPU baseline), never as the product path.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
outputs of the reference itself, run in the build container by ``tests/golden/make_golden.py``
(unmodified reference sources from /root/reference + the import shims in ``oracle/ref_shims``); the
resulting fixtures are committed under ``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.
"""
