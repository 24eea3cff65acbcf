"""ORACLE — test infrastructure, not product code.

CPU restatements of the reference's algorithm for the GRU/BiLSTM hot path plus the tooling that pins them to
the reference. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from here; the product (icassp2022-depression_b200/) never does and has no CPU fallback.
"""
