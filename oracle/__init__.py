"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see maxk_oracle.c)."""
