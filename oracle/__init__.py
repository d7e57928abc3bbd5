"""CPU oracle of the PFEMFort implicit hot path: TEST INFRASTRUCTURE ONLY (see pfem_oracle.c header)."""
