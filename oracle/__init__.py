"""CPU parity oracle for the PTM hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs; the
product package (rti_b200) never imports it."""
