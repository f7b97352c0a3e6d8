"""TEST INFRASTRUCTURE ONLY: the CPU oracle of the EdgeCalculator / FindNextOverlaps path.

Nothing in ``haploconduct_b200`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
