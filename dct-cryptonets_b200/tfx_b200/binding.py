"""ctypes binding of libtfx_b200.so (include/tfx.h).

PyTorch is used only for device memory (int64 tensors hold torus words bit-for-bit), streams and
torch.distributed; every arithmetic step of the encrypted path runs in the CUDA library.  There is no CPU
fallback: if the shared library or a CUDA device is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import weakref
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFX_LIB", os.path.join(_HERE, "libtfx_b200.so"))    # TFX_LIB: experiment builds only


class TfxError(RuntimeError):
    pass


class PbsParamsC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("k", C.c_uint32), ("N", C.c_uint32),
                ("bsk_base_log", C.c_uint32), ("bsk_level", C.c_uint32),
                ("ksk_base_log", C.c_uint32), ("ksk_level", C.c_uint32),
                ("reserved", C.c_uint32), ("lwe_std", C.c_double), ("glwe_std", C.c_double)]


@dataclass(frozen=True)
class PbsParams:
    """One PBS flavour (see tfx_pbs_params in include/tfx.h)."""
    n: int
    k: int
    N: int
    bsk_base_log: int
    bsk_level: int
    ksk_base_log: int
    ksk_level: int
    lwe_std: float
    glwe_std: float

    def to_c(self) -> PbsParamsC:
        return PbsParamsC(self.n, self.k, self.N, self.bsk_base_log, self.bsk_level, self.ksk_base_log,
                          self.ksk_level, 0, self.lwe_std, self.glwe_std)

    @property
    def big_dim(self) -> int:
        """dimension of the LWE ciphertexts this set's sample extraction produces (k*N); a keyset's big key may be longer"""
        return self.k * self.N


# every symbol include/tfx.h declares: name -> (restype, argtypes)
_VP, _U32, _U64, _I32, _I64, _SZ, _D = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int64, C.c_size_t, C.c_double
SYMBOLS = {
    "tfx_last_error": (C.c_char_p, []),
    "tfx_version": (C.c_char_p, []),
    "tfx_pbs_supported": (_I32, [_U32, _U32]),
    "tfx_ctx_create": (_I32, [_I32, _VP, _I32, C.POINTER(_VP)]),
    "tfx_ctx_destroy": (None, [_VP]),
    "tfx_ctx_set_stream": (_I32, [_VP, _VP]),
    "tfx_ctx_synchronize": (_I32, [_VP]),
    "tfx_keyset_generate": (_I32, [_VP, _U32, C.POINTER(PbsParamsC), _U32, _VP, _I32, C.POINTER(_VP)]),
    "tfx_keyset_create_empty": (_I32, [_VP, _U32, C.POINTER(PbsParamsC), _U32, C.POINTER(_VP)]),
    "tfx_keyset_destroy": (None, [_VP]),
    "tfx_keyset_drop_secret": (_I32, [_VP]),
    "tfx_keyset_get_secret": (_I32, [_VP, _I32, _VP]),
    "tfx_keyset_set_secret": (_I32, [_VP, _I32, _VP]),
    "tfx_keyset_get_ksk": (_I32, [_VP, _U32, _VP]),
    "tfx_keyset_set_ksk": (_I32, [_VP, _U32, _VP]),
    "tfx_keyset_get_bsk_fourier": (_I32, [_VP, _U32, _VP]),
    "tfx_keyset_set_bsk_fourier": (_I32, [_VP, _U32, _VP]),
    "tfx_keyset_get_bsk_standard": (_I32, [_VP, _U32, _VP]),
    "tfx_keyset_device_bytes": (_SZ, [_VP]),
    "tfx_lwe_encrypt": (_I32, [_VP, _VP, _I32, _D, _VP, _SZ, _VP, _U64, _VP]),
    "tfx_lwe_phase": (_I32, [_VP, _VP, _I32, _VP, _SZ, _VP]),
    "tfx_keyswitch_batch": (_I32, [_VP, _VP, _U32, _VP, _VP, _SZ, _U32, _U64]),
    "tfx_pbs_batch": (_I32, [_VP, _VP, _U32, _VP, _VP, _VP, _VP, _SZ, _I32, _U64]),
    "tfx_linear_conv2d": (_I32, [_VP, _VP, _U32, _U32, _U32, _U32, _VP, _U32, _U32, _U32, _U32, _U32, _VP, _U32, _U32, _U32, _VP]),
    "tfx_linear_axpby": (_I32, [_VP, _VP, _I64, _VP, _I64, _U64, _SZ, _U32, _VP]),
    "tfx_fft_tables": (_I32, [_U32, _VP]),
    "tfx_fft_forward": (_I32, [_VP, _U32, _VP, _SZ, _VP]),
    "tfx_fft_inverse": (_I32, [_VP, _U32, _VP, _SZ, _VP]),
    "tfx_probe_rate": (_I32, [_VP, _I32, C.POINTER(C.c_double)]),
    "tfx_launch_count": (_U64, []),
}

_lib = None


def load_library() -> C.CDLL:
    """Load libtfx_b200.so and bind every declared symbol.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TfxError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                           "(make -C dct-cryptonets_b200/csrc); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise TfxError(f"{what} failed ({rc}): {load_library().tfx_last_error().decode()}")


def seed_bytes(seed) -> bytes:
    if isinstance(seed, (bytes, bytearray)):
        b = bytes(seed)
    else:
        b = int(seed).to_bytes(16, "little")
    if len(b) != 16:
        raise ValueError("seed must be 16 bytes")
    return b


def _seed_buf(seed):
    return C.create_string_buffer(seed_bytes(seed), 16)


def _dptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensor must be CUDA and contiguous"
    return C.c_void_p(t.data_ptr())


def launch_count() -> int:
    return int(load_library().tfx_launch_count())


def fft_tables(N: int):
    tw = np.empty((N // 2, 2), dtype=np.float64)
    _check(load_library().tfx_fft_tables(N, tw.ctypes.data_as(C.c_void_p)), "tfx_fft_tables")
    return tw


class Context:
    """A device context bound to torch's current stream on `device`."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise TfxError("CUDA device required: tfx_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream if use_torch_stream else 0
        h = C.c_void_p()
        _check(self.lib.tfx_ctx_create(device, C.c_void_p(stream), 0 if use_torch_stream else 1, C.byref(h)), "tfx_ctx_create")
        self.h = h
        self._keysets = weakref.WeakSet()

    def set_stream(self, stream: torch.cuda.Stream):
        _check(self.lib.tfx_ctx_set_stream(self.h, C.c_void_p(stream.cuda_stream)), "tfx_ctx_set_stream")

    def synchronize(self):
        _check(self.lib.tfx_ctx_synchronize(self.h), "tfx_ctx_synchronize")

    def close(self):
        if getattr(self, "h", None):
            for ks in list(getattr(self, "_keysets", ())):       # key sets hold a pointer to the context: release them first
                ks.close()
            self.lib.tfx_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def probe_rate(self, which: int) -> float:
        """0: FP64 FMA FLOP/s, 1: u64 += u32*u64 MAC/s (measured on this device)"""
        r = C.c_double()
        _check(self.lib.tfx_probe_rate(self.h, which, C.byref(r)), "tfx_probe_rate")
        return r.value

    # ---- tensor helpers -------------------------------------------------------------------------
    def empty_u64(self, *shape) -> torch.Tensor:
        return torch.empty(*shape, dtype=torch.int64, device=self.device)

    def to_device_u64(self, a: np.ndarray) -> torch.Tensor:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).to(self.device)

    @staticmethod
    def to_host_u64(t: torch.Tensor) -> np.ndarray:
        return t.detach().cpu().numpy().view(np.uint64)

    # ---- test hooks -----------------------------------------------------------------------------
    def fft_forward(self, polys: torch.Tensor) -> torch.Tensor:
        P, N = polys.shape
        out = torch.empty(P, N // 2, 2, dtype=torch.float64, device=self.device)
        _check(self.lib.tfx_fft_forward(self.h, N, _dptr(polys), P, _dptr(out)), "tfx_fft_forward")
        return out

    def fft_inverse(self, freq: torch.Tensor) -> torch.Tensor:
        P, M, _ = freq.shape
        out = self.empty_u64(P, 2 * M)
        _check(self.lib.tfx_fft_inverse(self.h, 2 * M, _dptr(freq), P, _dptr(out)), "tfx_fft_inverse")
        return out

    # ---- leveled ops ----------------------------------------------------------------------------
    def conv2d(self, x: torch.Tensor, w: torch.Tensor, stride: int, pad: int, bias_pt: Optional[torch.Tensor] = None,
               oc_range: Optional[Sequence[int]] = None, depthwise: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        Cin, H, W, words = x.shape
        Cout, Cin_w, kh, kw = w.shape
        assert w.dtype == torch.int32 and Cin_w == (1 if depthwise else Cin)
        ob, oe = (0, Cout) if oc_range is None else oc_range
        Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        if out is None:
            out = self.empty_u64(oe - ob, Ho, Wo, words)
        assert tuple(out.shape) == (oe - ob, Ho, Wo, words)
        _check(self.lib.tfx_linear_conv2d(self.h, _dptr(x), Cin, H, W, words, _dptr(w), Cout, kh, kw, stride, pad,
                                          _dptr(bias_pt), ob, oe, int(depthwise), _dptr(out)), "tfx_linear_conv2d")
        return out

    def axpby(self, a: torch.Tensor, sa: int, b: Optional[torch.Tensor] = None, sb: int = 0, body_const: int = 0,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        words = a.shape[-1]
        count = a.numel() // words
        if out is None:
            out = torch.empty_like(a)
        _check(self.lib.tfx_linear_axpby(self.h, _dptr(a), sa, _dptr(b), sb, body_const & (2**64 - 1), count, words, _dptr(out)),
               "tfx_linear_axpby")
        return out


class KeySet:
    """Device-resident keys for one big LWE key and a list of PBS parameter sets."""

    def __init__(self, ctx: Context, params: Sequence[PbsParams], handle):
        self.ctx, self.params, self.h = ctx, list(params), handle
        ctx._keysets.add(self)
        self.big_dim = max(p.big_dim for p in params)           # the big LWE key; every set's GLWE key is a prefix of it

    # -- construction -------------------------------------------------------------------------------
    @staticmethod
    def _carr(params):
        arr = (PbsParamsC * len(params))(*[p.to_c() for p in params])
        return arr

    @classmethod
    def generate(cls, ctx: Context, params: Sequence[PbsParams], seed, keep_standard_bsk: bool = False) -> "KeySet":
        h = C.c_void_p()
        _check(ctx.lib.tfx_keyset_generate(ctx.h, max(p.big_dim for p in params), cls._carr(params), len(params), _seed_buf(seed),
                                           int(keep_standard_bsk), C.byref(h)), "tfx_keyset_generate")
        return cls(ctx, params, h)

    @classmethod
    def empty(cls, ctx: Context, params: Sequence[PbsParams]) -> "KeySet":
        h = C.c_void_p()
        _check(ctx.lib.tfx_keyset_create_empty(ctx.h, max(p.big_dim for p in params), cls._carr(params), len(params), C.byref(h)),
               "tfx_keyset_create_empty")
        return cls(ctx, params, h)

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.tfx_keyset_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(self.ctx.lib.tfx_keyset_device_bytes(self.h))

    # -- import / export (host numpy) -----------------------------------------------------------------
    def drop_secret(self):
        _check(self.ctx.lib.tfx_keyset_drop_secret(self.h), "tfx_keyset_drop_secret")

    def get_secret(self, set_id: int = -1) -> np.ndarray:
        dim = self.big_dim if set_id < 0 else self.params[set_id].n
        out = np.empty(dim, dtype=np.uint64)
        _check(self.ctx.lib.tfx_keyset_get_secret(self.h, set_id, out.ctypes.data_as(C.c_void_p)), "tfx_keyset_get_secret")
        return out

    def set_secret(self, key: np.ndarray, set_id: int = -1):
        key = np.ascontiguousarray(key, dtype=np.uint64)
        _check(self.ctx.lib.tfx_keyset_set_secret(self.h, set_id, key.ctypes.data_as(C.c_void_p)), "tfx_keyset_set_secret")

    def get_ksk(self, set_id: int) -> np.ndarray:
        p = self.params[set_id]
        out = np.empty((self.big_dim, p.ksk_level, p.n + 1), dtype=np.uint64)
        _check(self.ctx.lib.tfx_keyset_get_ksk(self.h, set_id, out.ctypes.data_as(C.c_void_p)), "tfx_keyset_get_ksk")
        return out

    def set_ksk(self, set_id: int, ksk: np.ndarray):
        p = self.params[set_id]
        ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
        assert ksk.shape == (self.big_dim, p.ksk_level, p.n + 1)
        _check(self.ctx.lib.tfx_keyset_set_ksk(self.h, set_id, ksk.ctypes.data_as(C.c_void_p)), "tfx_keyset_set_ksk")

    def _bsk_shape(self, set_id):
        p = self.params[set_id]
        return (p.n, p.k + 1, p.bsk_level, p.k + 1)

    def get_bsk_fourier(self, set_id: int) -> np.ndarray:
        p = self.params[set_id]
        out = np.empty(self._bsk_shape(set_id) + (p.N // 2, 2), dtype=np.float64)
        _check(self.ctx.lib.tfx_keyset_get_bsk_fourier(self.h, set_id, out.ctypes.data_as(C.c_void_p)), "tfx_keyset_get_bsk_fourier")
        return out

    def set_bsk_fourier(self, set_id: int, bsk: np.ndarray):
        p = self.params[set_id]
        bsk = np.ascontiguousarray(bsk, dtype=np.float64)
        assert bsk.shape == self._bsk_shape(set_id) + (p.N // 2, 2)
        _check(self.ctx.lib.tfx_keyset_set_bsk_fourier(self.h, set_id, bsk.ctypes.data_as(C.c_void_p)), "tfx_keyset_set_bsk_fourier")

    def get_bsk_standard(self, set_id: int) -> np.ndarray:
        p = self.params[set_id]
        out = np.empty(self._bsk_shape(set_id) + (p.N,), dtype=np.uint64)
        _check(self.ctx.lib.tfx_keyset_get_bsk_standard(self.h, set_id, out.ctypes.data_as(C.c_void_p)), "tfx_keyset_get_bsk_standard")
        return out

    # -- client ops ---------------------------------------------------------------------------------
    def encrypt(self, plaintexts: torch.Tensor, std: float, enc_seed, first_index: int = 0, key_sel: int = -1) -> torch.Tensor:
        dim = self.big_dim if key_sel < 0 else self.params[key_sel].n
        pts = plaintexts.contiguous().view(-1)
        out = self.ctx.empty_u64(pts.numel(), dim + 1)
        _check(self.ctx.lib.tfx_lwe_encrypt(self.ctx.h, self.h, key_sel, std, _dptr(pts), pts.numel(), _seed_buf(enc_seed),
                                            first_index, _dptr(out)), "tfx_lwe_encrypt")
        return out

    def phase(self, cts: torch.Tensor, key_sel: int = -1) -> torch.Tensor:
        dim = self.big_dim if key_sel < 0 else self.params[key_sel].n
        cts = cts.contiguous().view(-1, dim + 1)
        out = self.ctx.empty_u64(cts.shape[0])
        _check(self.ctx.lib.tfx_lwe_phase(self.ctx.h, self.h, key_sel, _dptr(cts), cts.shape[0], _dptr(out)), "tfx_lwe_phase")
        return out

    # -- server ops ---------------------------------------------------------------------------------
    # `ctx`: the context (stream + scratch) to launch on; defaults to the one the keys were made with.  A second context
    # on the same device lets two independent batches run on two streams (executor.py, multi-GPU shards).
    def keyswitch(self, set_id: int, cts: torch.Tensor, shift: int = 0, body_offset: int = 0,
                  out: Optional[torch.Tensor] = None, ctx: Optional[Context] = None) -> torch.Tensor:
        p = self.params[set_id]
        ctx = ctx if ctx is not None else self.ctx
        cts = cts.contiguous().view(-1, self.big_dim + 1)
        B = cts.shape[0]
        if out is None:
            out = ctx.empty_u64(B, p.n + 1)
        _check(ctx.lib.tfx_keyswitch_batch(ctx.h, self.h, set_id, _dptr(cts), _dptr(out), B, shift,
                                           body_offset & (2**64 - 1)), "tfx_keyswitch_batch")
        return out

    def pbs(self, set_id: int, cts: torch.Tensor, luts: torch.Tensor, lut_index: torch.Tensor, mode: int = 0,
            body_const: int = 0, out: Optional[torch.Tensor] = None, ctx: Optional[Context] = None) -> torch.Tensor:
        p = self.params[set_id]
        ctx = ctx if ctx is not None else self.ctx
        cts = cts.contiguous().view(-1, p.n + 1)
        B = cts.shape[0]
        assert luts.shape[-1] == p.N and lut_index.dtype == torch.int32 and lut_index.numel() == B
        if out is None:
            assert mode == 0
            out = ctx.empty_u64(B, self.big_dim + 1)
        assert out.numel() == B * (self.big_dim + 1) and out.is_contiguous()
        _check(ctx.lib.tfx_pbs_batch(ctx.h, self.h, set_id, _dptr(cts), _dptr(luts), _dptr(lut_index), _dptr(out), B,
                                     mode, body_const & (2**64 - 1)), "tfx_pbs_batch")
        return out
