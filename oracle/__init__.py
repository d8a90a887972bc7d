"""CPU checkers (test infrastructure).  See oracle/ref_oracle.c and oracle/Makefile."""
