"""oracle/ -- CPU restatement of the reference's LUT-qGEMM path.  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by flute_b200/.  See flute_oracle.py (numpy) and flute_oracle.c (C, OpenMP).
"""
