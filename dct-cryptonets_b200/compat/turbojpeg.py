"""Import-only stub (see README.md): reference data/cvtransforms.py:82-84 instantiates TurboJPEG() even on the 4x4 path."""
TJPF_RGB, TJPF_BGR, TJSAMP_420 = 0, 1, 2


class TurboJPEG:
    def __init__(self, *args, **kwargs):
        pass

    def encode(self, *args, **kwargs):
        raise RuntimeError("turbojpeg is not installed; the 8x8 JPEG path is provided by tfx_b200.dct_preprocess (filter_size=8)")

    decode = encode
