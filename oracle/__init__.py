"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker
or the timed CPU baseline.  ``horizonnet_b200`` never imports it (tests/test_boundary.py checks).
"""
