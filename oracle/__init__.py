"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (parity unpinned, see tph_dense.py header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
