"""CPU oracle for the AADG hot path — TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from aadg_b200/.  See each module's header for what pins it.
"""
