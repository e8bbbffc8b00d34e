"""CPU oracle for the GS-decoder hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package (see gs_oracle.c for the full notice,
including the "parity unpinned" statement for stages 2-4).
"""
