"""CPU oracle (test infrastructure).  See oracle/dbga_oracle.py for the contract:
only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
