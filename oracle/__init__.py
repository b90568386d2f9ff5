"""CPU oracle for the PointNet++ sem-seg attack hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is imported by the product package ``pointsecguard_b200``.  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it, as
the checker or as the timed CPU baseline, never as a fallback.
"""
