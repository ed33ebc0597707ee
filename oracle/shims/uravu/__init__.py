"""Stand-in for the absent `uravu` package (test infrastructure, see ../README.md)."""
