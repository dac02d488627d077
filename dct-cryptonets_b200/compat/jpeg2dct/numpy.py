"""Import-only stub (see ../README.md)."""


def loads(*args, **kwargs):
    raise RuntimeError("jpeg2dct is not installed; the 8x8 JPEG path is provided by tfx_b200.dct_preprocess (filter_size=8)")


load = loads
