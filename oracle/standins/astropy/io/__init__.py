"""stand-in package (test infrastructure)."""
