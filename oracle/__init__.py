"""CPU parity oracle (test infrastructure only; see oracle/oracle.py)."""
