"""CPU oracle of the xKV hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import this package, and only as the checker or the timed CPU baseline.  The product path
(``xkv_b200``) never imports it and has no CPU fallback.
"""
