"""Synthetic systems of the benchmark configurations and the duck-typed ``Options``
the tests and the bench pass to ``Mapping`` / ``BondSet``.  Benchmark and test infrastructure: not part of
the ``pycgtool_b200`` product package."""
