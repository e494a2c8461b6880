"""Test-only CPU checkers (see fabe13_oracle.c).  Not part of the product."""
