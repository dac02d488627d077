"""Drop-in for the slice of the Concrete / Concrete-ML Python API that the reference calls
(reference dct-cryptonets/homomorphic_eval.py:22-23): same import paths, same call signatures, executed by the
B200-native backend in tfx_b200 instead of concrete-python's CPU runtime."""
