"""ORACLE: CPU restatements of the reference's hot path, used only as the checker.

Test infrastructure.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs - never by the product package catfish_b200.
"""
