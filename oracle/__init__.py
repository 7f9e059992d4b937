"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see kl_oracle.c header).
Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs."""
