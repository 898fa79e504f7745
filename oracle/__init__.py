"""CPU oracle of the hot path — test infrastructure only (see mg_oracle.py header)."""
