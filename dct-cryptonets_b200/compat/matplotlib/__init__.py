"""Import-only stub (see ../README.md): plotting is out of scope."""
