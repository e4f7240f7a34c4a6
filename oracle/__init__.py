"""TEST INFRASTRUCTURE ONLY -- CPU/GPU oracles for the hot path.

Nothing under ``impersonator_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()``, and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
