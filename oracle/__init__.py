"""TEST INFRASTRUCTURE ONLY — CPU oracle for the PAIF fusion hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or as the reported CPU
baseline.  ``paif_b200`` never imports this package.
"""
