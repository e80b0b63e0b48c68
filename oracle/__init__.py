"""
oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the BaryonForge runner hot path (the reference is pure Python, so the
restatement is numpy/scipy plus a small C library for the un-vendored HEALPix boundary).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from here; the product package `baryonforge_b200` never does.

Parity status: the runner restatement (oracle/runners_port.py) is pinned against outputs of
the reference's own runner code executed in the build container (tests/golden/*.npz, made by
oracle/make_golden.py).  The HEALPix boundary (healpy, third party, absent, version unpinned)
is pinned only against healpy's docstring known-answer values and brute-force geometry:
"parity unpinned" at that boundary in the strict sense.
"""
