"""CPU oracle for the MAGI hot path: test infrastructure only (see magi_oracle.py header)."""
