"""Minimal stand-in for fairscale (test infrastructure; see ../README.md)."""
