"""Integer circuit IR + the compile front-end that builds it from a torch module.

This is our realisation of what the reference obtains from Concrete-ML's compile_torch_model /
compile_brevitas_qat_model (reference homomorphic_eval.py:276-295; intent summarised in SURVEY.md A.8):
  * the input is quantised in the clear to n_bits (symmetric);
  * every Conv2d becomes an integer conv with clear integer weights, the residual add and the average pool are
    leveled integer ops;
  * everything univariate between two integer tensors (BatchNorm, ReLU, quantisers, the 1/k^2 of the pool,
    re-quantisation) is fused into ONE per-channel table, evaluated after exact rounding of the accumulator to
    rounding_threshold_bits;
  * accumulator widths come from the calibration set (like Concrete's inputset-driven bit-width assignment).
The IR is plain data (numpy arrays + ints) so that the CUDA executor (executor.py), the clear integer evaluator
below and the CPU oracle's evaluator (oracle/circuit_oracle.py, test infrastructure) all run the same circuit.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.fx as fx
import torch.nn as nn
import torch.nn.functional as F

from .params import CircuitNoiseSpec, RoundedLookup


# --------------------------------------------------------------------------------------------------------
# IR
# --------------------------------------------------------------------------------------------------------
@dataclass
class QuantInfo:
    scale: float
    qmin: int
    qmax: int


@dataclass
class ConvOp:
    name: str
    src: int
    dst: int
    weight: np.ndarray          # int32 [Cout][Cin or 1][kh][kw], already multiplied by 2^lshift
    stride: int
    pad: int
    depthwise: bool
    in_shape: Tuple[int, int, int]
    out_shape: Tuple[int, int, int]
    acc_bits: int = 0           # w (one width per tensor, like Concrete's per-tensor bit-width assignment)
    offset: object = 0          # int64 [Cout]: added per output channel so the accumulator is unsigned: u = acc + offset in [0, 2^w)
    lshift: int = 0
    raw_weight: Optional[np.ndarray] = None   # before the encoding shift
    chan_bits: Optional[np.ndarray] = None    # per_channel_widths: int64 [Cout], width w_c <= acc_bits of every channel (None: all acc_bits)
    kind: str = "conv"


@dataclass
class AddOp:
    name: str
    a: int
    b: int
    dst: int
    shape: Tuple[int, int, int]
    acc_bits: int = 0
    offset: object = 0          # int64 [C], see ConvOp.offset
    sa: int = 1                 # 2^lshift of operand a
    sb: int = 1
    chan_bits: Optional[np.ndarray] = None    # see ConvOp.chan_bits
    kind: str = "add"


@dataclass
class FusedAddOp:
    """dst = a + m_c * b (fuse_residual, opt-in): `a` is the accumulator of a conv whose float chain is affine per channel (BatchNorm),
    `b` a quantised tensor (the identity shortcut).  real(a-chain) + real(b) = g_c * (a + r_c * b) + beta_c with r_c = s_b / g_c; the
    integer m_c = round(r_c) stands in for r_c (|r_c| >= 8, so the shortcut's weight is off by < 1/16 and typically ~1 %), and
    the block needs one table lookup instead of two.  The conv is emitted at this op's width (same encoding), so the sum is leveled."""
    name: str
    a: int                      # conv accumulator (value id)
    b: int                      # quantised tensor (value id)
    dst: int
    shape: Tuple[int, int, int]
    m: np.ndarray               # int64 [C]
    acc_bits: int = 0
    offset: object = 0          # int64 [C]
    sb: Optional[np.ndarray] = None   # int64 [C]: m_c * 2^(emitted width of b - acc_bits), set when widths are final
    chan_bits: Optional[np.ndarray] = None
    kind: str = "fadd"


@dataclass
class LinOp:
    """dst = sum_i coef_i * Tap_i(src_i): a leveled linear combination of quantised tensors that share one scale.  A term with a
    tap (ky, kx) reads element (y * stride + ky - pad, x * stride + kx - pad) of its source (zero outside the image); a term with
    ky = -1 is already at the output resolution.  Building block of MaxPool2d, reference models/backbone.py:153-160,252-259:
    the running maximum is m_j = m_{j-1} + relu(t_j - m_{j-1}), i.e. r_j = relu(t_j - t_0 - r_1 - ... - r_{j-1}) as one lookup on
    a LinOp per window element and max = t_0 + r_1 + ... + r_{k*k-1} as a last LinOp (the sources are >= 0, so the zero padding of
    a tap equals MaxPool2d's -inf padding)."""
    name: str
    terms: list                 # [[source value id, coefficient, ky, kx], ...]; the first term carries a tap
    dst: int
    in_shape: Tuple[int, int, int]
    shape: Tuple[int, int, int]
    kernel: int
    stride: int
    pad: int
    acc_bits: int = 0
    offset: object = 0          # int64 [C], see ConvOp.offset
    shifts: Optional[list] = None   # per term: log2 of the encoding shift (emitted width of the source - acc_bits)
    chan_bits: Optional[np.ndarray] = None
    kind: str = "lin"

    def eff_coefs(self) -> List[int]:
        return [int(c) << int(sh) for (_, c, _, _), sh in zip(self.terms, self.shifts or [0] * len(self.terms))]


@dataclass
class TluOp:
    name: str
    src: int                    # accumulator value (output of a conv / add)
    dst: int
    shape: Tuple[int, int, int]
    acc_bits: int               # w
    keep_bits: int              # t' = min(w, rounding bits)
    tables: np.ndarray          # int64 [C][2^keep_bits] integer outputs q
    out: QuantInfo
    out_width: int = 0          # encoding width of the produced ciphertexts (delta = 2^(63 - out_width))
    chan_bits: Optional[np.ndarray] = None   # per-channel widths of the source accumulator (None: acc_bits for every channel)
    kind: str = "tlu"

    @property
    def lsbs(self) -> int:
        """bits removed from the widest channel (every channel when chan_bits is None)"""
        return self.acc_bits - self.keep_bits

    def chan_widths(self) -> np.ndarray:
        C = self.tables.shape[0]
        return np.full(C, self.acc_bits, dtype=np.int64) if self.chan_bits is None else np.asarray(self.chan_bits, dtype=np.int64)

    def chan_lsbs(self) -> np.ndarray:
        """int64 [C]: bits the exact rounding removes per channel.  All channels share the encoding 2^(63 - acc_bits); a channel of
        width w_c < acc_bits simply has fewer low bits to extract, and its table lookup reads the ciphertext scaled by 2^(acc_bits - w_c)."""
        return np.maximum(self.chan_widths() - self.keep_bits, 0)


@dataclass
class Circuit:
    input_shape: Tuple[int, int, int]
    input_quant: QuantInfo
    input_id: int
    input_width: int
    ops: List[object]
    output_id: int
    output_shape: Tuple[int, ...]
    output_is_acc: bool         # True: decrypt an accumulator (offset/width below); False: a TLU output
    output_width: int
    output_offset: object       # int64 [C] (per channel of the output accumulator) or 0
    output_scale: float         # real value = output_scale * integer
    n_bits: int
    rounding_bits: int
    p_error: float
    rounding_method: str = "exact"   # "exact": bit-extraction chain; "approximate": the table lookup is applied
                                     # to the unrounded accumulator, low bits ride inside the LUT box (reference README.md:96-113)

    # ---- statistics ------------------------------------------------------------------------------------
    def lookups(self) -> List[TluOp]:
        return [op for op in self.ops if op.kind == "tlu"]

    def pbs_count(self) -> Dict[str, int]:
        tlu = sum(int(np.prod(op.shape)) for op in self.lookups())
        bit = (sum(int(np.prod(op.shape[1:])) * int(op.chan_lsbs().sum()) for op in self.lookups())
               if self.rounding_method == "exact" else 0)
        return {"tlu": tlu, "bit": bit, "total": tlu + bit}

    def macs(self) -> int:
        total = 0
        for op in self.ops:
            if op.kind == "conv":
                cout, ho, wo = op.out_shape
                total += cout * ho * wo * int(np.prod(op.weight.shape[1:]))
        return total

    def maximum_integer_bit_width(self) -> int:
        widths = [op.acc_bits for op in self.ops if op.kind in ("conv", "add", "fadd", "lin")]
        return max(widths + [self.input_width])

    def noise_spec(self, input_std: float = 2.0 ** -50) -> CircuitNoiseSpec:
        producers = {op.dst: op for op in self.ops}
        looks = []
        for op in self.lookups():
            lin = producers[op.src]
            if lin.kind == "conv":
                w = lin.weight.astype(np.float64)
                norm2 = float((w.reshape(w.shape[0], -1) ** 2).sum(axis=1).max())
                fresh = lin.src == self.input_id
            elif lin.kind == "lin":
                norm2 = float(sum(c * c for c in lin.eff_coefs()))
                fresh = False
            elif lin.kind == "fadd":
                w = producers[lin.a].weight.astype(np.float64)
                norm2 = float((w.reshape(w.shape[0], -1) ** 2).sum(axis=1).max()) + float((np.asarray(lin.sb, dtype=np.float64) ** 2).max())
                fresh = False
            else:
                norm2 = float(lin.sa ** 2 + lin.sb ** 2)
                fresh = False
            shift = int(op.acc_bits - op.chan_widths().min())         # narrowest channel: its table lookup amplifies the noise by 2^shift
            if self.rounding_method == "exact":
                looks.append(RoundedLookup(op.acc_bits, op.keep_bits, max(norm2, 1.0), fresh, int(np.prod(op.shape)), shift))
            else:
                # approximate: one t-bit lookup straight on the accumulator; the accumulator noise counts at the table's
                # granularity (the low bits are signal that shifts the rounding threshold, not noise)
                looks.append(RoundedLookup(op.keep_bits, op.keep_bits, max(norm2, 1.0), fresh, int(np.prod(op.shape)), shift))
        return CircuitNoiseSpec(looks, self.p_error, input_std)

    def to_text(self) -> str:
        """Printable listing (plays the role of fhe_circuit.mlir, reference homomorphic_eval.py:311)."""
        lines = [f"circuit(input %{self.input_id}: eint<{self.input_width}>[{','.join(map(str, self.input_shape))}], "
                 f"n_bits={self.n_bits}, rounding={self.rounding_bits} ({self.rounding_method}), p_error={self.p_error})"]
        for op in self.ops:
            if op.kind == "conv":
                k = "sum_pool" if op.depthwise else "conv2d"
                lines.append(f"  %{op.dst} = {k}(%{op.src}) {{weight=i32{list(op.weight.shape)}, stride={op.stride}, pad={op.pad}, "
                             f"lshift={op.lshift}, offset={_offset_text(op.offset)}}} : eint<{op.acc_bits}>{list(op.out_shape)}   // {op.name}")
            elif op.kind == "fadd":
                lines.append(f"  %{op.dst} = fused_add(%{op.a}, %{op.b} * [{int(np.min(op.sb))}..{int(np.max(op.sb))}]/channel) "
                             f"{{offset={_offset_text(op.offset)}}} : eint<{op.acc_bits}>{list(op.shape)}   // {op.name}")
            elif op.kind == "lin":
                tt = " + ".join(f"{c} * %{sv}" + (f"[tap {ky},{kx}]" if ky >= 0 else "") for (sv, _, ky, kx), c in zip(op.terms, op.eff_coefs()))
                lines.append(f"  %{op.dst} = lincomb({tt}) {{window={op.kernel}, stride={op.stride}, pad={op.pad}, offset={_offset_text(op.offset)}}} "
                             f": eint<{op.acc_bits}>{list(op.shape)}   // {op.name}")
            elif op.kind == "add":
                lines.append(f"  %{op.dst} = add(%{op.a} * {op.sa}, %{op.b} * {op.sb}) {{offset={_offset_text(op.offset)}}} : eint<{op.acc_bits}>{list(op.shape)}   // {op.name}")
            else:
                l_c = op.chan_lsbs()
                lsbs_txt = str(op.lsbs) if l_c.min() == l_c.max() else f"{int(l_c.min())}..{int(l_c.max())}/channel"
                lines.append(f"  %{op.dst} = round_lsbs<{lsbs_txt}>.table_lookup(%{op.src}) {{tables=i64{list(op.tables.shape)}}} : "
                             f"eint<{op.out_width}>{list(op.shape)}   // {op.name}")
        lines.append(f"  return %{self.output_id}")
        return "\n".join(lines)


# --------------------------------------------------------------------------------------------------------
# Clear integer evaluator (exact semantics of the circuit; also the noise-free model behind fhe='simulate')
# --------------------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------------------
# Serialisation: JSON header + raw arrays (no pickle: a bundle comes from the model provider and is opened by the
# process that holds the secret key).  Unknown op kinds or fields are rejected.
# --------------------------------------------------------------------------------------------------------
_OP_CLASSES = {"conv": ConvOp, "add": AddOp, "fadd": FusedAddOp, "lin": LinOp, "tlu": TluOp}
_TUPLE_FIELDS = {"in_shape", "out_shape", "shape", "input_shape", "output_shape"}


def _enc_value(v, arrays: list):
    if isinstance(v, np.ndarray):
        arrays.append(np.ascontiguousarray(v))
        return {"__array__": len(arrays) - 1}
    if isinstance(v, QuantInfo):
        return {"__quant__": [float(v.scale), int(v.qmin), int(v.qmax)]}
    if isinstance(v, (tuple, list)):
        return [_enc_value(x, arrays) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if v is None or isinstance(v, (bool, int, float, str)):
        return v
    raise TypeError(f"cannot serialise a {type(v).__name__} in a circuit")


def _dec_value(v, arrays: list, name: str):
    if isinstance(v, dict):
        if set(v) == {"__array__"}:
            return np.array(arrays[int(v["__array__"])])          # own, writable copy
        if set(v) == {"__quant__"}:
            sc, lo, hi = v["__quant__"]
            return QuantInfo(float(sc), int(lo), int(hi))
        raise ValueError(f"unknown object in circuit field {name!r}")
    if isinstance(v, list):
        seq = [_dec_value(x, arrays, name) for x in v]
        return tuple(seq) if name in _TUPLE_FIELDS else seq
    return v


def circuit_to_portable(circ: "Circuit"):
    """-> (JSON-able header, list of numpy arrays)"""
    import dataclasses
    arrays: list = []
    head = {}
    for f in dataclasses.fields(Circuit):
        if f.name == "ops":
            continue
        head[f.name] = _enc_value(getattr(circ, f.name), arrays)
    ops = []
    for op in circ.ops:
        d = {"kind": op.kind}
        for f in dataclasses.fields(type(op)):
            if f.name != "kind":
                d[f.name] = _enc_value(getattr(op, f.name), arrays)
        ops.append(d)
    head["ops"] = ops
    return head, arrays


def circuit_from_portable(head: dict, arrays: list) -> "Circuit":
    import dataclasses
    head = dict(head)
    ops = []
    for d in head.pop("ops"):
        d = dict(d)
        cls = _OP_CLASSES.get(d.pop("kind", None))
        if cls is None:
            raise ValueError("unknown op kind in circuit bundle")
        names = {f.name for f in dataclasses.fields(cls)} - {"kind"}
        if set(d) - names:
            raise ValueError(f"unknown fields in {cls.__name__}: {sorted(set(d) - names)}")
        ops.append(cls(**{k: _dec_value(v, arrays, k) for k, v in d.items()}))
    names = {f.name for f in dataclasses.fields(Circuit)} - {"ops"}
    if set(head) - names:
        raise ValueError(f"unknown fields in Circuit: {sorted(set(head) - names)}")
    return Circuit(ops=ops, **{k: _dec_value(v, arrays, k) for k, v in head.items()})


def _int_conv(x: np.ndarray, op: ConvOp, weight: np.ndarray) -> np.ndarray:
    """x: int64 [B][C][H][W] -> int64; float64 conv is exact for these magnitudes (< 2^53)."""
    xt = torch.from_numpy(x.astype(np.float64))
    wt = torch.from_numpy(weight.astype(np.float64))
    groups = x.shape[1] if op.depthwise else 1
    y = F.conv2d(xt, wt, stride=op.stride, padding=op.pad, groups=groups)
    return np.rint(y.numpy()).astype(np.int64)


def window_tap(x: np.ndarray, ky: int, kx: int, kernel: int, stride: int, pad: int) -> np.ndarray:
    """x int64 [B][C][H][W] -> [B][C][Ho][Wo]: element (y*stride + ky - pad, x*stride + kx - pad), zero outside"""
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 * pad - kernel) // stride + 1, (W + 2 * pad - kernel) // stride + 1
    xp = np.zeros((B, C, H + 2 * pad, W + 2 * pad), dtype=x.dtype)
    xp[:, :, pad:pad + H, pad:pad + W] = x
    return np.ascontiguousarray(xp[:, :, ky:ky + (Ho - 1) * stride + 1:stride, kx:kx + (Wo - 1) * stride + 1:stride])


def lin_apply(op: "LinOp", vals: dict) -> np.ndarray:
    out = None
    for sv, coef, ky, kx in op.terms:
        v = vals[sv] if ky < 0 else window_tap(vals[sv], ky, kx, op.kernel, op.stride, op.pad)
        out = coef * v if out is None else out + coef * v
    return out


def channel_offsets(offset, channels: int) -> np.ndarray:
    """per-channel accumulator offsets as int64 [C] (a scalar means the same offset for every channel)"""
    o = np.asarray(offset, dtype=np.int64)
    return np.full(channels, int(o), dtype=np.int64) if o.ndim == 0 else o


def _bcast_offset(offset):
    o = np.asarray(offset, dtype=np.int64)
    return o.reshape(1, -1, 1, 1) if o.ndim == 1 else o


def _offset_text(offset) -> str:
    o = np.asarray(offset).reshape(-1)
    return str(int(o[0])) if o.min() == o.max() else f"[{int(o.min())}..{int(o.max())}]/channel"


def quantize_input(circ: Circuit, x: np.ndarray) -> np.ndarray:
    q = np.rint(np.asarray(x, dtype=np.float64) / circ.input_quant.scale)
    return np.clip(q, circ.input_quant.qmin, circ.input_quant.qmax).astype(np.int64)


def _chan_view(a: np.ndarray) -> np.ndarray:
    return np.asarray(a, dtype=np.int64).reshape(1, -1, 1, 1)


def tlu_apply(op: TluOp, offset, acc: np.ndarray) -> np.ndarray:
    """acc int64 [B][C][H][W] (true accumulator, before offset) -> q int64; models the padding-bit wrap.
    offset: scalar or int64 [C].  Widths / removed bits may differ per channel (TluOp.chan_bits)."""
    t = op.keep_bits
    w = _chan_view(op.chan_widths())
    lsbs = _chan_view(op.chan_lsbs())
    half = np.where(lsbs > 0, np.left_shift(1, np.maximum(lsbs - 1, 0)), 0)
    u = (acc + _bcast_offset(offset) + half) & (np.left_shift(1, w + 1) - 1)
    idx = np.right_shift(u, lsbs)                    # in [0, 2^(t+1))
    neg = idx >= (1 << t)
    idx = idx & ((1 << t) - 1)
    C = op.tables.shape[0]
    ch = np.arange(C).reshape(1, C, 1, 1)
    q = op.tables[np.broadcast_to(ch, idx.shape), idx]
    return np.where(neg, -q, q)


def tlu_apply_noisy(op: TluOp, offset, acc: np.ndarray, exact: bool, norm2: float, fresh: bool, nm: "NoiseModel",
                    rng: np.random.Generator) -> np.ndarray:
    """Like tlu_apply, but every PBS decision sees the modelled ciphertext noise: the accumulator noise
    (weights x PBS output noise), the keyswitch + mod-switch noise at each PBS input and the output noise of every
    extracted bit.  All quantities in units of the accumulator LSB (delta_w = 2^-(w+1) of the torus, w = the tensor's width)."""
    W, t = op.acc_bits, op.keep_bits
    w_c = _chan_view(op.chan_widths()).astype(np.float64)
    lsbs_c = _chan_view(op.chan_lsbs())
    lsb = 2.0 ** -(W + 1)                                            # torus fraction of one accumulator unit
    half = np.where((lsbs_c > 0) & exact, np.left_shift(1, np.maximum(lsbs_c - 1, 0)), 0)
    v_src = nm.input_var if fresh else nm.var_tlu_out
    x = (acc + _bcast_offset(offset) + half).astype(np.float64) + rng.normal(0.0, np.sqrt(norm2 * v_src) / lsb, size=acc.shape)
    if exact:
        for b in range(int(lsbs_c.max())):
            # phase of (x << (W - b)) + 1/4 turn, in turns; bit b is its top bit; only channels that still have bits to remove
            ph = x * 2.0 ** (W - b) * lsb + 0.25 + rng.normal(0.0, np.sqrt(nm.var_bit_in), size=acc.shape)
            bit = (np.floor(ph * 2.0) % 2.0)
            x = np.where(lsbs_c > b, x - bit * (1 << b) + rng.normal(0.0, np.sqrt(nm.var_bit_out) / lsb, size=acc.shape), x)
    # table lookup on t bits + padding of the ciphertext scaled by 2^(W - w_c): index = round(phase / delta_t)
    ph = x * lsb * 2.0 ** (W - w_c) + rng.normal(0.0, np.sqrt(nm.var_tlu_in), size=acc.shape)
    idx = np.floor(ph * 2.0 ** (t + 1) + 0.5).astype(np.int64) & ((1 << (t + 1)) - 1)
    neg = idx >= (1 << t)
    idx = idx & ((1 << t) - 1)
    C = op.tables.shape[0]
    ch = np.arange(C).reshape(1, C, 1, 1)
    q = op.tables[np.broadcast_to(ch, idx.shape), idx]
    return np.where(neg, -q, q)


@dataclass
class NoiseModel:
    """variances (torus units) of the quantities a PBS decision depends on; built from the picked parameter sets"""
    var_tlu_out: float
    var_bit_out: float
    var_tlu_in: float     # keyswitch + mod-switch of the table set
    var_bit_in: float     # keyswitch + mod-switch of the bit-extraction set
    input_var: float

    @classmethod
    def from_params(cls, tlu, bit, input_std: float) -> "NoiseModel":
        from . import params as P
        big = max(tlu.k * tlu.N, bit.k * bit.N)
        return cls(P.var_pbs_out(tlu), P.var_pbs_out(bit), P.var_keyswitch(tlu, big) + P.var_modswitch(tlu),
                   P.var_keyswitch(bit, big) + P.var_modswitch(bit), input_std ** 2)


def evaluate_clear(circ: Circuit, q_in: np.ndarray, collect: Optional[dict] = None, noise: Optional[NoiseModel] = None,
                   rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """q_in int64 [B][C][H][W] -> integer outputs [B][...] (accumulator or table outputs).
    noise=None: exact integer semantics.  noise=NoiseModel: Monte-Carlo of the encrypted run (fhe='simulate')."""
    vals = {circ.input_id: q_in.astype(np.int64)}
    offs = {}
    spec = {op.dst: lk for op, lk in zip(circ.lookups(), circ.noise_spec().lookups)} if noise is not None else {}
    if noise is not None and rng is None:
        rng = np.random.default_rng(0)
    for op in circ.ops:
        if op.kind == "conv":
            vals[op.dst] = _int_conv(vals[op.src], op, op.raw_weight if op.raw_weight is not None else op.weight)
            offs[op.dst] = op.offset
        elif op.kind == "add":
            vals[op.dst] = vals[op.a] + vals[op.b]
            offs[op.dst] = op.offset
        elif op.kind == "fadd":
            vals[op.dst] = vals[op.a] + _chan_view(op.m) * vals[op.b]
            offs[op.dst] = op.offset
        elif op.kind == "lin":
            vals[op.dst] = lin_apply(op, vals)
            offs[op.dst] = op.offset
        elif noise is None:
            vals[op.dst] = tlu_apply(op, offs[op.src], vals[op.src])
        else:
            lk = spec[op.dst]
            vals[op.dst] = tlu_apply_noisy(op, offs[op.src], vals[op.src], circ.rounding_method == "exact", lk.weight_norm2,
                                           lk.fresh_inputs, noise, rng)
        if collect is not None:
            collect[op.dst] = vals[op.dst]
    out = vals[circ.output_id]
    return out.reshape(out.shape[0], *circ.output_shape)


def dequantize_output(circ: Circuit, q_out: np.ndarray) -> np.ndarray:
    return q_out.astype(np.float64) * circ.output_scale


# --------------------------------------------------------------------------------------------------------
# Front-end
# --------------------------------------------------------------------------------------------------------
class _Tracer(fx.Tracer):
    """Quant* modules (Brevitas or the stubs) and standard layers are leaves."""

    def is_leaf_module(self, m: nn.Module, qualname: str) -> bool:
        if type(m).__name__.startswith("Quant"):
            return True
        return super().is_leaf_module(m, qualname)


@dataclass
class _Sym:
    """a float tensor = chain(scale * integer accumulator) not yet turned into a table"""
    lin: int                                  # value id of the integer tensor underneath
    lin_is_acc: bool                          # True: conv/add output; False: a quantised tensor (input / table output)
    scale: float                              # real = scale * int (before the chain)
    chain: List[Callable[[torch.Tensor], torch.Tensor]]
    shape: Tuple[int, ...]
    affine_only: bool = True                  # chain is a pure positive scaling (can be decoded without a table)
    affine_factor: float = 1.0
    quant: Optional[QuantInfo] = None         # the chain ends in a QAT activation quantiser: values lie on this grid
    pending_input: bool = False               # the raw float input, not quantised yet (its first consumer decides how)


# head-room (fraction of its span, per side) kept by a channel that is narrower than its tensor; 1/16 brings the rate of
# out-of-range accumulators on unseen inputs back to what tensor-wide calibration gives
NARROW_SLACK = float(os.environ.get("TFX_NARROW_SLACK", "0.0625"))


def _bits_for_range(lo: int, hi: int) -> int:
    span = hi - lo + 1
    return max(1, int(math.ceil(math.log2(span)))) if span > 1 else 1


def _weight_bits(mod: nn.Module, n_bits: int) -> int:
    for attr in ("weight_bit_width", "bit_width"):
        v = getattr(mod, attr, None)
        if isinstance(v, int):
            return v
    return n_bits


def _act_bits(mod: nn.Module, n_bits: int) -> int:
    v = getattr(mod, "act_bit_width", None)
    return v if isinstance(v, int) else n_bits


def _act_quant_spec(mod: nn.Module) -> Optional[QuantInfo]:
    """activation quantiser of a QAT module (QuantIdentity / QuantReLU, reference models/backbone.py:71-73,83,231,249,261,278):
    scale and integer range as the module defines them (learned threshold), not calibrated.  Reads the compat shim
    (compat/brevitas) or Brevitas' own accessors; None means 'no quantiser information' (treated like the float op)."""
    if hasattr(mod, "tfx_act_quant"):
        scale, lo, hi = mod.tfx_act_quant()
        return QuantInfo(float(scale), int(lo), int(hi))
    try:                                                       # Brevitas proper (untested here: not installable)
        scale = float(mod.quant_act_scale())
        bits = int(mod.quant_act_bit_width())
        signed = bool(mod.is_quant_act_signed)
        narrow = bool(getattr(mod, "is_quant_act_narrow_range", False))
    except Exception:
        return None
    if signed:
        return QuantInfo(scale, -(1 << (bits - 1)) + (1 if narrow else 0), (1 << (bits - 1)) - 1)
    return QuantInfo(scale, 0, (1 << bits) - 1 - (1 if narrow else 0))


def _weight_quant_spec(mod: nn.Module):
    """(int32 weights, scale) of a QuantConv2d as the module defines them, or None"""
    if hasattr(mod, "tfx_weight_quant"):
        w, scale = mod.tfx_weight_quant()
        return w.detach().cpu().numpy().astype(np.int32), float(scale)
    try:                                                       # Brevitas proper (untested here)
        qw = mod.quant_weight()
        return qw.int().detach().cpu().numpy().astype(np.int32), float(qw.scale)
    except Exception:
        return None


def _fake_quant_fn(q: QuantInfo):
    return lambda y: torch.clamp(torch.round(y / q.scale), q.qmin, q.qmax) * q.scale


class CircuitBuilder:
    def __init__(self, model: nn.Module, calib: torch.Tensor, n_bits: int = 5, rounding_threshold_bits: int = 6,
                 p_error: float = 0.01, range_margin: float = 0.0, per_channel_offsets: bool = True,
                 per_channel_widths: Optional[bool] = None, fuse_residual: bool = False):
        self.model = model.eval()
        # opt-in (TFX_FUSE_RESIDUAL=1 for A/B runs): validated against the oracle on the CPU only, never run on a GPU yet
        self.fuse_residual = fuse_residual or os.environ.get("TFX_FUSE_RESIDUAL", "0") == "1"
        if os.environ.get("TFX_PER_CHANNEL_OFFSETS", "1") == "0":   # A/B knob: the Concrete-like tensor-wide layout
            per_channel_offsets, per_channel_widths = False, False
        if per_channel_widths is None:                       # default on; TFX_PER_CHANNEL_WIDTHS=0 switches it off for A/B measurements
            per_channel_widths = per_channel_offsets and os.environ.get("TFX_PER_CHANNEL_WIDTHS", "1") != "0"
        if per_channel_widths and not per_channel_offsets:
            raise ValueError("per_channel_widths needs per_channel_offsets")
        self.per_channel_offsets, self.per_channel_widths = per_channel_offsets, per_channel_widths
        self.n_bits, self.t, self.p_error, self.margin = n_bits, rounding_threshold_bits, p_error, range_margin
        self.calib = calib.detach().to(torch.float64).cpu()
        self.ops: List[object] = []
        self.next_id = 0
        self.ints: Dict[int, np.ndarray] = {}          # calibration integers per value id
        self.qinfo: Dict[int, QuantInfo] = {}          # for quantised tensors
        self.acc_of: Dict[int, object] = {}            # value id -> producing linear op
        self.materialized: Dict[int, int] = {}         # fx node id -> quantised value id
        self.consumers: Dict[int, List[Tuple[object, str]]] = {}

    def _new(self) -> int:
        self.next_id += 1
        return self.next_id - 1

    # ---- accumulator range / width -----------------------------------------------------------------------
    def _finish_acc(self, op, acc: np.ndarray):
        """One width per tensor.  per_channel_offsets (default): every channel is centred in [0, 2^w) on its own, so the width
        only has to cover the widest channel span (times 1 + 2 * range_margin) instead of the union of all channel ranges —
        often one bit (= one bit-extraction PBS per element) less — and every channel keeps equal slack below and above
        (a channel aligned at 0 wraps below zero at the first noisy or out-of-calibration value).
        per_channel_offsets=False: one tensor-wide offset, u = acc - min over the tensor."""
        lo_c = acc.min(axis=(0, 2, 3)).astype(np.int64)
        hi_c = acc.max(axis=(0, 2, 3)).astype(np.int64)
        if not self.per_channel_offsets:
            lo_c[:], hi_c[:] = lo_c.min(), hi_c.max()
        span_c = hi_c - lo_c
        span = int(span_c.max())

        def width_for(sp: int) -> int:
            need = sp + 2 * int(math.ceil(self.margin * max(1, sp)))
            w_ = _bits_for_range(0, need)
            while w_ > self.t and need + (1 << (w_ - self.t - 1)) >= (1 << w_):       # the rounded index must still fit
                w_ += 1
            return w_

        w = width_for(span)
        op.acc_bits = w
        if self.per_channel_offsets:
            if self.per_channel_widths and w > self.t:
                # a channel narrower than the tensor keeps NARROW_SLACK of its span as head-room on each side (the widest
                # channels, which set the tensor's width, keep what tensor-wide calibration gives them)
                w_c = np.array([min(w, max(self.t, width_for(int(math.ceil(sp * (1.0 + 2.0 * NARROW_SLACK))))))
                                for sp in span_c], dtype=np.int64)
                if (w_c < w).any():
                    op.chan_bits = w_c
            else:
                w_c = np.full(span_c.shape, w, dtype=np.int64)
            half_c = np.where(w_c > self.t, np.left_shift(1, np.maximum(w_c - self.t - 1, 0)), 0)
            slack_c = (np.left_shift(1, w_c) - half_c - 1 - span_c) // 2             # centre every channel: equal slack below and above
            op.offset = slack_c - lo_c
        else:
            m = int(math.ceil(self.margin * max(1, span)))
            op.offset = np.full(lo_c.shape, m - int(lo_c[0]), dtype=np.int64)

    # ---- table construction ------------------------------------------------------------------------------
    def _materialize(self, node_key, sym: _Sym, forced_scale: Optional[float] = None, out_bits: Optional[int] = None) -> int:
        if node_key in self.materialized:
            vid = self.materialized[node_key]
            if forced_scale is not None and abs(self.qinfo[vid].scale - forced_scale) > 1e-12 * forced_scale:
                raise NotImplementedError("residual add of two already-quantised tensors with different scales")
            return vid
        if not sym.lin_is_acc:
            if sym.chain:
                raise NotImplementedError("univariate op directly on a quantised tensor without a linear op in between")
            self.materialized[node_key] = sym.lin
            return sym.lin
        lin_op = self.acc_of[sym.lin]
        acc = self.ints[sym.lin]                                             # [B][C][H][W]
        w, off = lin_op.acc_bits, lin_op.offset
        keep = min(w, self.t)
        C = acc.shape[1]
        w_c = np.full(C, w, dtype=np.int64) if lin_op.chan_bits is None else np.asarray(lin_op.chan_bits, dtype=np.int64)
        lsbs = np.maximum(w_c - keep, 0)                                     # [C] bits removed per channel
        # float function per channel on every representable rounded accumulator value
        idx = np.arange(1 << keep, dtype=np.int64)
        off = channel_offsets(off, C)
        acc_vals = np.left_shift(idx[None, :], lsbs[:, None]) - off[:, None]  # [C][2^keep]: value the index stands for, per channel
        xin = torch.from_numpy((acc_vals.astype(np.float64) * sym.scale).reshape(1, C, 1 << keep, 1))
        y = xin
        for fn in sym.chain:
            y = fn(y)
        y = y.reshape(C, 1 << keep).numpy()
        # output quantiser: calibrate on the values the calibration set actually reaches
        half = np.where(lsbs > 0, np.left_shift(1, np.maximum(lsbs - 1, 0)), 0)
        cal_idx = np.clip(np.right_shift(acc + (off + half).reshape(1, C, 1, 1), lsbs.reshape(1, C, 1, 1)), 0, (1 << keep) - 1)
        ch = np.arange(C).reshape(1, C, 1, 1)
        y_cal = y[np.broadcast_to(ch, cal_idx.shape), cal_idx]
        nb = out_bits if out_bits is not None else self.n_bits
        signed = bool(y_cal.min() < 0)
        qmax = (1 << (nb - 1)) - 1 if signed else (1 << nb) - 1
        if forced_scale is None and sym.quant is not None:
            # QAT: the module's own quantiser is the table's output grid (y already lies on it)
            scale, qlo, qhi = sym.quant.scale, sym.quant.qmin, sym.quant.qmax
        elif forced_scale is None:
            amax = float(np.abs(y_cal).max())
            scale = (amax / qmax) if amax > 0 else 1.0
            qlo, qhi = (-qmax if signed else 0), qmax
        else:
            scale = forced_scale
            reach = int(math.ceil(float(np.abs(y_cal).max()) / scale)) + 1
            qlo, qhi = (-reach if signed else 0), reach
        tables = np.clip(np.rint(y / scale), qlo, qhi).astype(np.int64)
        vid = self._new()
        op = TluOp(f"tlu_{vid}", sym.lin, vid, tuple(acc.shape[1:]), w, keep, tables, QuantInfo(scale, qlo, qhi),
                   chan_bits=lin_op.chan_bits)
        self.ops.append(op)
        self.ints[vid] = tlu_apply(op, off, acc)
        self.qinfo[vid] = op.out
        self.materialized[node_key] = vid
        return vid

    # ---- graph walk -------------------------------------------------------------------------------------
    def build(self) -> Circuit:
        tracer = _Tracer()
        graph = tracer.trace(self.model)
        mods = dict(self.model.named_modules())
        env: Dict[fx.Node, _Sym] = {}
        out_sym = None
        for node in graph.nodes:
            if node.op == "placeholder":
                env[node] = _Sym(-1, False, 1.0, [], tuple(self.calib.shape[1:]), pending_input=True)
            elif node.op == "call_module":
                m = mods[node.target]
                tname = type(m).__name__
                aq = _act_quant_spec(m) if tname in ("QuantIdentity", "QuantReLU") else None
                if env[node.args[0]].pending_input:
                    # the input quantiser: a leading QuantIdentity defines it (QAT), otherwise n_bits symmetric from the calibration set
                    env[node.args[0]] = self._quantize_input(aq if tname == "QuantIdentity" else None)
                    if tname == "QuantIdentity" and aq is not None:
                        env[node] = env[node.args[0]]
                        continue
                src = env[node.args[0]]
                if isinstance(m, nn.Conv2d) or tname == "QuantConv2d":
                    env[node] = self._conv(node, m, src, node.args[0])
                elif isinstance(m, nn.BatchNorm2d):
                    env[node] = self._chain(src, _bn_fn(m), False)
                elif aq is not None:
                    if not src.lin_is_acc:
                        raise NotImplementedError("activation quantiser directly on a quantised tensor (needs a linear op first)")
                    fn = _fake_quant_fn(aq) if tname == "QuantIdentity" else (lambda y, f=_fake_quant_fn(aq): f(torch.relu(y)))
                    env[node] = self._chain(src, fn, False)
                    env[node].quant = aq
                elif isinstance(m, nn.ReLU) or tname == "QuantReLU":
                    env[node] = self._chain(src, torch.relu, False)
                elif tname == "QuantIdentity" or isinstance(m, (nn.Identity, nn.Dropout)):
                    env[node] = src if not src.lin_is_acc else self._chain(src, _identity, src.affine_only, keep_affine=True)
                elif isinstance(m, nn.AvgPool2d):
                    env[node] = self._avgpool(node, m, src, node.args[0])
                elif isinstance(m, nn.MaxPool2d):
                    env[node] = self._maxpool(node, m, src, node.args[0])
                elif isinstance(m, nn.Flatten):
                    env[node] = _Sym(src.lin, src.lin_is_acc, src.scale, src.chain, (int(np.prod(src.shape)),), src.affine_only, src.affine_factor,
                                     src.quant)
                else:
                    raise NotImplementedError(f"unsupported module {tname} at {node.target}")
            elif node.op == "call_function":
                for arg in node.args:
                    if isinstance(arg, fx.Node) and env[arg].pending_input:
                        env[arg] = self._quantize_input(None)
                fname = getattr(node.target, "__name__", str(node.target))
                if fname in ("add", "iadd"):
                    env[node] = self._add(node, env[node.args[0]], env[node.args[1]], node.args[0], node.args[1])
                elif fname == "relu":
                    env[node] = self._chain(env[node.args[0]], torch.relu, False)
                elif fname == "flatten":
                    src = env[node.args[0]]
                    env[node] = _Sym(src.lin, src.lin_is_acc, src.scale, src.chain, (int(np.prod(src.shape)),), src.affine_only, src.affine_factor,
                                     src.quant)
                else:
                    raise NotImplementedError(f"unsupported function {fname}")
            elif node.op == "output":
                out_sym = env[node.args[0]]
                out_key = node.args[0]
            else:
                raise NotImplementedError(f"unsupported fx node {node.op}")
        return self._finish(out_sym, out_key)

    def _quantize_input(self, spec: Optional[QuantInfo]) -> _Sym:
        x = self.calib
        if spec is None:
            amax = float(x.abs().max())
            qmax = (1 << (self.n_bits - 1)) - 1
            spec = QuantInfo(amax / qmax if amax > 0 else 1.0, -qmax, qmax)
        vid = self._new()
        self.input_id = vid
        self.input_quant = spec
        self.ints[vid] = np.clip(np.rint(x.numpy() / spec.scale), spec.qmin, spec.qmax).astype(np.int64)
        self.qinfo[vid] = spec
        return _Sym(vid, False, spec.scale, [], tuple(x.shape[1:]))

    def _chain(self, src: _Sym, fn, affine: bool, keep_affine: bool = False) -> _Sym:
        if not src.lin_is_acc:
            raise NotImplementedError("univariate op on a quantised tensor (needs a linear op first)")
        return _Sym(src.lin, True, src.scale, src.chain + [fn], src.shape,
                    src.affine_only and (affine or keep_affine), src.affine_factor)

    def _register(self, vid: int, op, role: str):
        self.consumers.setdefault(vid, []).append((op, role))

    def _conv(self, node, m, src: _Sym, src_key) -> _Sym:
        q = self._materialize(src_key, src)
        wspec = _weight_quant_spec(m) if type(m).__name__ == "QuantConv2d" else None
        if wspec is not None:                                                # QAT: the module's own weight quantiser
            wq, s_w = wspec
        else:
            wb = _weight_bits(m, self.n_bits)
            wq_max = (1 << (wb - 1)) - 1
            wf = m.weight.detach().to(torch.float64)
            s_w = float(wf.abs().max()) / wq_max if float(wf.abs().max()) > 0 else 1.0
            wq = torch.clamp(torch.round(wf / s_w), -wq_max, wq_max).numpy().astype(np.int32)
        if getattr(m, "groups", 1) != 1:
            raise NotImplementedError("grouped convolution")
        if m.bias is not None:
            raise NotImplementedError("conv bias (the reference models use bias=False)")
        stride = m.stride[0]; pad = m.padding[0]
        cin, h, w_ = self.ints[q].shape[1:]
        vid = self._new()
        op = ConvOp(f"conv_{node.target}", q, vid, wq, stride, pad, False, (cin, h, w_), (0, 0, 0), raw_weight=wq.copy())
        acc = _int_conv(self.ints[q], op, wq)
        op.out_shape = tuple(acc.shape[1:])
        self._finish_acc(op, acc)
        self.ops.append(op)
        self.ints[vid] = acc
        self.acc_of[vid] = op
        self._register(q, op, "src")
        return _Sym(vid, True, self.qinfo[q].scale * s_w, [], op.out_shape)

    def _avgpool(self, node, m, src: _Sym, src_key) -> _Sym:
        q = self._materialize(src_key, src)
        k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
        s = m.stride if isinstance(m.stride, int) else m.stride[0]
        c, h, w_ = self.ints[q].shape[1:]
        wq = np.ones((c, 1, k, k), dtype=np.int32)
        vid = self._new()
        op = ConvOp(f"sumpool_{node.target}", q, vid, wq, s, 0, True, (c, h, w_), (0, 0, 0), raw_weight=wq.copy())
        acc = _int_conv(self.ints[q], op, wq)
        op.out_shape = tuple(acc.shape[1:])
        self._finish_acc(op, acc)
        self.ops.append(op)
        self.ints[vid] = acc
        self.acc_of[vid] = op
        self._register(q, op, "src")
        f = 1.0 / (k * k)
        return _Sym(vid, True, self.qinfo[q].scale, [lambda y, f=f: y * f], op.out_shape, True, f)

    def _lin(self, name: str, terms: list, in_shape, out_shape, k: int, s: int, p: int) -> "LinOp":
        vid = self._new()
        op = LinOp(name, [list(map(int, t)) for t in terms], vid, tuple(in_shape), tuple(out_shape), k, s, p, shifts=[0] * len(terms))
        acc = lin_apply(op, self.ints)
        self._finish_acc(op, acc)
        self.ops.append(op)
        self.ints[vid] = acc
        self.acc_of[vid] = op
        for idx, t in enumerate(op.terms):
            self._register(t[0], op, f"t{idx}")
        return op

    def _maxpool(self, node, m, src: _Sym, src_key) -> _Sym:
        """MaxPool2d on a quantised, non-negative tensor (the reference puts it after the stem ReLU, backbone.py:153-160): a chain of
        k*k - 1 lookups r_j = relu(t_j - running maximum) over the window elements t_j, all on the source's scale (see LinOp)."""
        q = self._materialize(src_key, src)
        scale = self.qinfo[q].scale
        one = lambda v: v if isinstance(v, int) else v[0]
        k, s, p = one(m.kernel_size), one(m.stride if m.stride is not None else m.kernel_size), one(m.padding)
        if one(m.dilation) != 1 or m.ceil_mode:
            raise NotImplementedError("MaxPool2d with dilation or ceil_mode")
        if p > 0 and self.ints[q].min() < 0:
            raise NotImplementedError("padded MaxPool2d on a tensor with negative values (a tap pads with zero, MaxPool2d with -inf)")
        c, h, w_ = self.ints[q].shape[1:]
        out_shape = (c, (h + 2 * p - k) // s + 1, (w_ + 2 * p - k) // s + 1)
        taps = [(ky, kx) for ky in range(k) for kx in range(k)]
        rs: List[int] = []
        for j in range(1, len(taps)):
            terms = [[q, 1, *taps[j]], [q, -1, *taps[0]]] + [[r, -1, -1, -1] for r in rs]
            d = self._lin(f"maxdiff{j}_{node.target}", terms, (c, h, w_), out_shape, k, s, p)
            rs.append(self._materialize(("maxpool", node.name, j), _Sym(d.dst, True, scale, [torch.relu], out_shape, False), forced_scale=scale))
        top = self._lin(f"maxpool_{node.target}", [[q, 1, *taps[0]]] + [[r, 1, -1, -1] for r in rs], (c, h, w_), out_shape, k, s, p)
        return _Sym(top.dst, True, scale, [], out_shape)

    def _try_fused_add(self, node, a: _Sym, b: _Sym, ka, kb) -> Optional[_Sym]:
        """a: unmaterialised conv accumulator with a per-channel affine chain (BatchNorm); b: quantised tensor.  See FusedAddOp."""
        qb = self.materialized.get(kb) if b.lin_is_acc else (b.lin if not b.chain else None)   # b must already exist as a quantised tensor
        if not (a.lin_is_acc and a.chain and ka not in self.materialized and qb is not None):
            return None
        conv = self.acc_of.get(a.lin)
        if conv is None or conv.kind != "conv" or conv.depthwise:
            return None
        C = self.ints[a.lin].shape[1]
        probe = torch.tensor([0.0, 1.0, 1024.0], dtype=torch.float64).reshape(1, 1, 3, 1).expand(1, C, 3, 1) * a.scale
        y = probe
        for fn in a.chain:
            y = fn(y)
        y = y.reshape(C, 3).numpy()
        g = y[:, 1] - y[:, 0]                                              # real value per accumulator unit, per channel
        if not np.allclose(y[:, 2], y[:, 0] + 1024.0 * g, rtol=1e-9, atol=1e-12) or not np.all(np.isfinite(g)) or np.any(g == 0):
            return None                                                    # chain is not affine
        r = self.qinfo[qb].scale / g
        if np.abs(r).min() < 8.0 or np.abs(r).max() > 4096.0:
            return None
        m = np.rint(r).astype(np.int64)
        vid = self._new()
        op = FusedAddOp(f"fadd_{node.name}", a.lin, qb, vid, tuple(self.ints[a.lin].shape[1:]), m)
        acc = self.ints[a.lin] + m.reshape(1, C, 1, 1) * self.ints[qb]
        self._finish_acc(op, acc)
        # the conv is emitted directly at the fused accumulator's encoding: no table lookup reads it, no offset of its own
        conv.acc_bits, conv.offset, conv.chan_bits = op.acc_bits, np.zeros(C, dtype=np.int64), None
        self.ops.append(op)
        self.ints[vid] = acc
        self.acc_of[vid] = op
        self._register(qb, op, "b")
        return _Sym(vid, True, a.scale, list(a.chain), op.shape)

    def _add(self, node, a: _Sym, b: _Sym, ka, kb) -> _Sym:
        if self.fuse_residual:
            fused = self._try_fused_add(node, a, b, ka, kb) or self._try_fused_add(node, b, a, kb, ka)
            if fused is not None:
                return fused
        # operands must share one scale: an already-quantised operand dictates it, otherwise the larger range does
        qa = self.materialized.get(ka) if a.lin_is_acc else a.lin
        qb = self.materialized.get(kb) if b.lin_is_acc else b.lin
        if qa is not None and qb is None:
            qb = self._materialize(kb, b, forced_scale=self.qinfo[qa].scale)
        elif qb is not None and qa is None:
            qa = self._materialize(ka, a, forced_scale=self.qinfo[qb].scale)
        elif qa is None and qb is None:
            qa = self._materialize(ka, a)
            qb = self._materialize(kb, b, forced_scale=self.qinfo[qa].scale)
        elif abs(self.qinfo[qa].scale - self.qinfo[qb].scale) > 1e-12 * self.qinfo[qa].scale:
            raise NotImplementedError("residual add of two already-quantised tensors with different scales")
        vid = self._new()
        op = AddOp(f"add_{node.name}", qa, qb, vid, tuple(self.ints[qa].shape[1:]))
        acc = self.ints[qa] + self.ints[qb]
        self._finish_acc(op, acc)
        self.ops.append(op)
        self.ints[vid] = acc
        self.acc_of[vid] = op
        self._register(qa, op, "a")
        self._register(qb, op, "b")
        return _Sym(vid, True, self.qinfo[qa].scale, [], op.shape)

    def _finish(self, out_sym: _Sym, out_key) -> Circuit:
        if out_sym.lin_is_acc and out_sym.affine_only:
            out_id, is_acc = out_sym.lin, True
            lin = self.acc_of[out_id]
            out_width, out_off = lin.acc_bits, lin.offset
            out_scale = out_sym.scale * out_sym.affine_factor
        else:
            vid = self._materialize(out_key, out_sym) if out_sym.lin_is_acc else out_sym.lin
            out_id, is_acc = vid, False
            qi = self.qinfo[vid]
            out_width = _bits_for_range(qi.qmin, qi.qmax) + 1
            out_off, out_scale = 0, qi.scale
        # encoding widths: a quantised tensor is emitted at the width of its widest consumer; narrower consumers shift
        width_of: Dict[int, int] = {}
        for vid, cons in self.consumers.items():
            width_of[vid] = max(op.acc_bits for op, _ in cons)
        for vid, cons in self.consumers.items():
            for op, role in cons:
                ls = width_of[vid] - op.acc_bits
                if op.kind == "conv":
                    op.lshift = ls
                    op.weight = (op.raw_weight.astype(np.int64) << ls).astype(np.int32)
                elif op.kind == "fadd":
                    op.sb = op.m << ls
                elif op.kind == "lin":
                    op.shifts[int(role[1:])] = ls
                elif role == "a":
                    op.sa = 1 << ls
                else:
                    op.sb = 1 << ls
        for op in self.ops:
            if op.kind == "tlu":
                op.out_width = width_of.get(op.dst, out_width)
        return Circuit(tuple(self.calib.shape[1:]), self.input_quant, self.input_id, width_of[self.input_id], self.ops, out_id,
                       tuple(out_sym.shape), is_acc, out_width, out_off, out_scale, self.n_bits, self.t, self.p_error)


def _identity(y):
    return y


def _bn_fn(m: nn.BatchNorm2d):
    mean = m.running_mean.detach().to(torch.float64).reshape(1, -1, 1, 1)
    var = m.running_var.detach().to(torch.float64).reshape(1, -1, 1, 1)
    g = (m.weight.detach().to(torch.float64) if m.affine else torch.ones_like(m.running_mean, dtype=torch.float64)).reshape(1, -1, 1, 1)
    b = (m.bias.detach().to(torch.float64) if m.affine else torch.zeros_like(m.running_mean, dtype=torch.float64)).reshape(1, -1, 1, 1)
    inv = g / torch.sqrt(var + m.eps)
    return lambda y: (y - mean) * inv + b


def build_circuit(model: nn.Module, calib: torch.Tensor, n_bits: int = 5, rounding_threshold_bits: int = 6,
                  p_error: float = 0.01, range_margin: float = 0.0, rounding_method: str = "exact",
                  per_channel_offsets: bool = True, per_channel_widths: Optional[bool] = None,
                  fuse_residual: bool = False) -> Circuit:
    if rounding_method not in ("exact", "approximate"):
        raise ValueError("rounding_method must be 'exact' or 'approximate'")
    circ = CircuitBuilder(model, calib, n_bits, rounding_threshold_bits, p_error, range_margin, per_channel_offsets,
                          per_channel_widths, fuse_residual).build()
    circ.rounding_method = rounding_method
    return circ
