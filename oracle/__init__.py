"""CPU oracle for the svgrasterize hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(``svgrasterize.py_b200``) never does; it fails loudly when its CUDA library is
missing instead of falling back to anything here.
"""
