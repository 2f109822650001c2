"""CPU oracle for the Telogator2 hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``telogator2_b200``)
never imports it and fails loudly without its CUDA extension.

See ``oracle/tg_oracle.c`` for the pinning status of each part (scan + RNG pinned
against the reference / CPython, the distance half against the reference's own code around a stand-in score();
only the Biopython NW score itself is PARITY UNPINNED).
"""
from .pyoracle import *  # noqa: F401,F403
