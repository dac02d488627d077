"""ctypes front for oracle/tfhe_oracle.c — TEST INFRASTRUCTURE ONLY (parity unpinned, see the C header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (dct-cryptonets_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtfhe_oracle.so")

u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tfhe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def _seed(seed) -> np.ndarray:
    if isinstance(seed, (bytes, bytearray)):
        b = bytes(seed)
    else:
        b = int(seed).to_bytes(16, "little")
    assert len(b) == 16
    return np.frombuffer(b, dtype=np.uint8).copy()


ST_BIGKEY, ST_SMALLKEY = 1, 2


def prf_fill(seed, stream: int, first: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint64)
    lib().orc_prf_fill(_p(_seed(seed)), C.c_uint64(stream), C.c_uint64(first), C.c_uint64(count), _p(out))
    return out


def gauss_fill(seed, stream: int, first: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.float64)
    lib().orc_gauss_fill(_p(_seed(seed)), C.c_uint64(stream), C.c_uint64(first), C.c_uint64(count), _p(out))
    return out


def gen_binary_key(seed, purpose: int, set_id: int, dim: int) -> np.ndarray:
    key = np.empty(dim, dtype=np.uint64)
    lib().orc_gen_binary_key(_p(_seed(seed)), C.c_int(purpose), C.c_uint32(set_id), C.c_uint32(dim), _p(key))
    return key


def lwe_encrypt(key, std: float, pts, seed, first_index: int = 0) -> np.ndarray:
    key = np.ascontiguousarray(key, dtype=np.uint64)
    pts = np.ascontiguousarray(pts, dtype=np.uint64).ravel()
    out = np.empty((pts.size, key.size + 1), dtype=np.uint64)
    lib().orc_lwe_encrypt(_p(key), C.c_uint32(key.size), C.c_double(std), _p(pts), C.c_uint64(pts.size),
                          _p(_seed(seed)), C.c_uint64(first_index), _p(out))
    return out


def lwe_phase(key, cts) -> np.ndarray:
    key = np.ascontiguousarray(key, dtype=np.uint64)
    cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, key.size + 1)
    out = np.empty(cts.shape[0], dtype=np.uint64)
    lib().orc_lwe_phase(_p(key), C.c_uint32(key.size), _p(cts), C.c_uint64(cts.shape[0]), _p(out))
    return out


def decompose(xs, base_log: int, level: int) -> np.ndarray:
    xs = np.ascontiguousarray(xs, dtype=np.uint64).ravel()
    out = np.empty((xs.size, level), dtype=np.int64)
    lib().orc_decompose(_p(xs), C.c_uint64(xs.size), C.c_int(base_log), C.c_int(level), _p(out))
    return out


def gen_ksk(big_key, small_key, base_log, level, std, seed, set_id=0) -> np.ndarray:
    big_key = np.ascontiguousarray(big_key, dtype=np.uint64)
    small_key = np.ascontiguousarray(small_key, dtype=np.uint64)
    out = np.empty((big_key.size, level, small_key.size + 1), dtype=np.uint64)
    lib().orc_gen_ksk(_p(big_key), C.c_uint32(big_key.size), _p(small_key), C.c_uint32(small_key.size),
                      C.c_int(base_log), C.c_int(level), C.c_double(std), _p(_seed(seed)), C.c_uint32(set_id), _p(out))
    return out


def keyswitch(ksk, cts, base_log, level, shift=0, body_offset=0) -> np.ndarray:
    big_dim, lv, n1 = ksk.shape
    assert lv == level
    cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, big_dim + 1)
    out = np.empty((cts.shape[0], n1), dtype=np.uint64)
    lib().orc_keyswitch(_p(ksk), C.c_uint32(big_dim), C.c_uint32(n1 - 1), C.c_int(base_log), C.c_int(level),
                        _p(cts), C.c_uint64(cts.shape[0]), C.c_int(shift), C.c_uint64(body_offset & (2**64 - 1)), _p(out))
    return out


def fft_tables(N: int):
    tw = np.empty((N // 2, 2), dtype=np.float64)
    lib().orc_fft_tables(C.c_uint32(N), _p(tw))
    return tw


def fft_forward(poly) -> np.ndarray:
    poly = np.ascontiguousarray(poly, dtype=np.float64)
    N = poly.size
    out = np.empty((N // 2, 2), dtype=np.float64)
    lib().orc_fft_forward(C.c_uint32(N), _p(poly), _p(out))
    return out


def fft_inverse(freq) -> np.ndarray:
    freq = np.ascontiguousarray(freq, dtype=np.float64)
    N = freq.shape[0] * 2
    out = np.empty(N, dtype=np.float64)
    lib().orc_fft_inverse(C.c_uint32(N), _p(freq), _p(out))
    return out


def double_to_torus(v) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.float64).ravel()
    out = np.empty(v.size, dtype=np.uint64)
    lib().orc_double_to_torus(_p(v), C.c_uint64(v.size), _p(out))
    return out


def poly_mul_negacyclic(a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_poly_mul_negacyclic(_p(a), _p(b), C.c_uint32(a.size), _p(out))
    return out


def gen_bsk(small_key, big_key, k, N, base_log, level, std, seed, set_id=0) -> np.ndarray:
    small_key = np.ascontiguousarray(small_key, dtype=np.uint64)
    big_key = np.ascontiguousarray(big_key, dtype=np.uint64)[: k * N]      # the GLWE key is the first k*N bits of the big key
    assert big_key.size == k * N
    n = small_key.size
    out = np.empty((n, k + 1, level, k + 1, N), dtype=np.uint64)
    lib().orc_gen_bsk(_p(small_key), C.c_uint32(n), _p(big_key), C.c_uint32(k), C.c_uint32(N), C.c_int(base_log),
                      C.c_int(level), C.c_double(std), _p(_seed(seed)), C.c_uint32(set_id), _p(out))
    return out


def bsk_to_fourier(bsk) -> np.ndarray:
    n, k1, level, _, N = bsk.shape
    out = np.empty((n, k1, level, k1, N // 2, 2), dtype=np.float64)
    lib().orc_bsk_to_fourier(_p(np.ascontiguousarray(bsk)), C.c_uint32(n), C.c_uint32(k1 - 1), C.c_uint32(N),
                             C.c_int(level), _p(out))
    return out


def pbs(bsk_f, base_log, cts, luts, lut_index, mode=0, body_const=0, out=None, big_dim=None) -> np.ndarray:
    n, k1, level, _, M, _ = bsk_f.shape
    k, N = k1 - 1, 2 * M
    big = k * N if big_dim is None else int(big_dim)
    assert big >= k * N
    cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, n + 1)
    luts = np.ascontiguousarray(luts, dtype=np.uint64).reshape(-1, N)
    lut_index = np.ascontiguousarray(lut_index, dtype=np.uint32).ravel()
    assert lut_index.size == cts.shape[0] and (lut_index.max(initial=0) < luts.shape[0])
    if out is None:
        assert mode == 0
        out = np.empty((cts.shape[0], big + 1), dtype=np.uint64)
    assert out.flags.c_contiguous and out.shape == (cts.shape[0], big + 1)
    lib().orc_pbs(_p(bsk_f), C.c_uint32(n), C.c_uint32(k), C.c_uint32(N), C.c_uint32(big), C.c_int(base_log), C.c_int(level),
                  _p(cts), _p(luts), _p(lut_index), C.c_uint64(cts.shape[0]), C.c_int(mode),
                  C.c_uint64(body_const & (2**64 - 1)), _p(out))
    return out


def conv2d(x, w, stride=1, pad=0, bias_pt=None, depthwise=False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint64)
    Cin, H, W, words = x.shape
    w = np.ascontiguousarray(w, dtype=np.int32)
    Cout, Cin2, kh, kw = w.shape
    assert Cin2 == (1 if depthwise else Cin)
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    out = np.empty((Cout, Ho, Wo, words), dtype=np.uint64)
    bp = None if bias_pt is None else np.ascontiguousarray(bias_pt, dtype=np.uint64)
    lib().orc_conv2d(_p(x), C.c_uint32(Cin), C.c_uint32(H), C.c_uint32(W), C.c_uint32(words), _p(w), C.c_uint32(Cout),
                     C.c_uint32(kh), C.c_uint32(kw), C.c_uint32(stride), C.c_uint32(pad), _p(bp), C.c_uint32(int(depthwise)), _p(out))
    return out


def axpby(a, sa: int, b=None, sb: int = 0, body_const: int = 0) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    words = a.shape[-1]
    count = a.size // words
    out = np.empty_like(a)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    lib().orc_axpby(_p(a), C.c_int64(sa), _p(bb), C.c_int64(sb), C.c_uint64(body_const & (2**64 - 1)),
                    C.c_uint64(count), C.c_uint32(words), _p(out))
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """size of the OpenMP team for every later call (bench.py: all host cores, whatever OMP_NUM_THREADS says)"""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()
