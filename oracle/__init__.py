"""CPU oracle for the BP4 hot path -- test infrastructure only (see bp4_oracle.py)."""
