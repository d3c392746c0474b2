"""CPU oracle for the PCGmix hot path — test infrastructure only (see pcgmix_oracle.py)."""
