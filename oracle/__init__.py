"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/sis_oracle.h)."""
