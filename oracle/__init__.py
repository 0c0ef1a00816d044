"""CPU oracle for the ingest hot path.  TEST INFRASTRUCTURE ONLY (see oracle/vt_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
