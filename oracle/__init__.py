"""CPU oracle (test infrastructure only) -- see oracle/gp_oracle.py."""
