"""CPU oracle for the NanoWrap hot path -- test infrastructure only (see nanowrap_oracle.py)."""
