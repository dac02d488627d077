"""Import-only stub (see README.md): the reference's summary() call is commented out (homomorphic_eval.py:232-246)."""


def summary(*args, **kwargs):
    raise RuntimeError("torchinfo is not installed")
