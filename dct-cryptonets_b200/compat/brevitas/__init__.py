"""Minimal stand-in for Brevitas (not installable here): only what reference models/backbone.py uses. See ../README.md."""
__version__ = "0.8.0+tfx_b200.compat"
