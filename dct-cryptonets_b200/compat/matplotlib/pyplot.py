"""Import-only stub (see ../README.md): plotting is out of scope."""


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    def _missing(*args, **kwargs):
        raise RuntimeError("matplotlib is not installed")
    return _missing
