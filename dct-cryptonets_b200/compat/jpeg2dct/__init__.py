"""Import-only stub (see ../README.md)."""
