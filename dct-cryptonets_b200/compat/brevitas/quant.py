"""Quantiser tags used by the reference (models/backbone.py:5,221,227).  In Brevitas these are injector classes; here they only
carry the defaults the layers read (signedness, narrow range, bit width)."""


class Int8WeightPerTensorFloat:
    bit_width = 8
    signed = True
    narrow_range = True
    per_channel = False


class Int8ActPerTensorFloat:
    bit_width = 8
    signed = True
    narrow_range = False


class Uint8ActPerTensorFloat:
    bit_width = 8
    signed = False
    narrow_range = False
