"""CPU oracle for the OrientedRepPoints dense-inference hot path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (orientedreppoints_b200) never
imports this package and fails loudly when its CUDA library is missing.
"""
