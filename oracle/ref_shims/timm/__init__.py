"""Minimal stand-in for timm (test infrastructure; see ../README.md)."""
