"""Stand-in for the absent `statsmodels` package (test infrastructure, see ../README.md)."""
