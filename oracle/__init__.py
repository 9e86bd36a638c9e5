"""TEST INFRASTRUCTURE (CPU oracles).  Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
