"""ORACLE package: CPU restatements of the reference path, test infrastructure only.

Nothing under ``mlff_distiller_b200`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` do.
"""
