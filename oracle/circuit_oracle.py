"""CPU evaluation of a compiled Circuit with the oracle's TFHE primitives — TEST INFRASTRUCTURE ONLY.

Independent of the CUDA executor: it walks the same IR (plain data produced by tfx_b200.circuit) but builds its own
LUT polynomials and calls only oracle/tfhe_oracle.c.  Used (i) by the parity tests to check the GPU circuit run
ciphertext word by ciphertext word, (ii) by bench.py's cpu_baseline / --impl reference legs as the timed CPU path.
PARITY UNPINNED (see tfhe_oracle.c header).
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np

from . import oracle as O

MASK64 = (1 << 64) - 1


class OracleKeys:
    """keys for a list of PBS parameter sets, generated with the oracle's own keygen (same seed => same keys as the GPU)"""

    def __init__(self, params, seed, with_bsk_sets=None):
        self.params = list(params)
        self.big_dim = max(p.k * p.N for p in params)
        self.big = O.gen_binary_key(seed, O.ST_BIGKEY, 0, self.big_dim)
        self.small, self.ksk, self.bsk_f = [], [], []
        for s, p in enumerate(params):
            small = O.gen_binary_key(seed, O.ST_SMALLKEY, s, p.n)
            self.small.append(small)
            self.ksk.append(O.gen_ksk(self.big, small, p.ksk_base_log, p.ksk_level, p.lwe_std, seed, s))
            if with_bsk_sets is None or s in with_bsk_sets:
                bsk = O.gen_bsk(small, self.big, p.k, p.N, p.bsk_base_log, p.bsk_level, p.glwe_std, seed, s)
                self.bsk_f.append(O.bsk_to_fourier(bsk))
            else:
                self.bsk_f.append(None)

    @classmethod
    def from_arrays(cls, params, big, small, ksk, bsk_f):
        self = cls.__new__(cls)
        self.params, self.big, self.small, self.ksk, self.bsk_f = list(params), big, list(small), list(ksk), list(bsk_f)
        return self


def lut_poly(table_row: np.ndarray, keep_bits: int, N: int, out_width: int) -> np.ndarray:
    size = 1 << keep_bits
    box = N // size
    out = np.empty(N, dtype=np.uint64)
    shift = 63 - out_width
    for j in range(N):
        slot = (j + box // 2) // box
        if slot < size:
            out[j] = (int(table_row[slot]) << shift) & MASK64
        else:
            out[j] = (-(int(table_row[0]) << shift)) & MASK64
    return out


def encrypt_input(circ, keys: OracleKeys, q_in: np.ndarray, std: float, enc_seed) -> np.ndarray:
    shift = 63 - circ.input_width
    pts = np.array([(int(v) << shift) & MASK64 for v in q_in.reshape(-1)], dtype=np.uint64)
    return O.lwe_encrypt(keys.big, std, pts, enc_seed)


def decrypt_output(circ, keys: OracleKeys, cts: np.ndarray) -> np.ndarray:
    ph = O.lwe_phase(keys.big, cts)
    w = circ.output_width
    out = np.empty(ph.size, dtype=np.int64)
    for i, p in enumerate(ph):
        u = ((int(p) + (1 << (62 - w))) >> (63 - w)) & ((1 << (w + 1)) - 1)
        if circ.output_is_acc:
            offs = np.asarray(circ.output_offset, dtype=np.int64).reshape(-1)          # scalar or one offset per output channel
            out[i] = u - int(offs[0] if offs.size == 1 else offs[i // (ph.size // offs.size)])
        else:
            out[i] = u - (1 << (w + 1)) if u >= (1 << w) else u
    return out


def _body_constants(offset, lsbs_c, acc_bits: int, channels: int) -> np.ndarray:
    """u64 [C]: (offset_c + half LSB of the rounding that follows, per channel) at the accumulator's encoding;
    offset is a scalar or one value per channel, lsbs_c None (no rounding follows) or int64 [C]"""
    offs = np.asarray(offset, dtype=np.int64).reshape(-1)
    if offs.size == 1:
        offs = np.repeat(offs, channels)
    ls = np.zeros(channels, dtype=np.int64) if lsbs_c is None else np.asarray(lsbs_c, dtype=np.int64)
    return np.array([((int(o) + ((1 << (int(l) - 1)) if l > 0 else 0)) << (63 - acc_bits)) & MASK64 for o, l in zip(offs, ls)],
                    dtype=np.uint64)


def run_circuit(circ, keys: OracleKeys, in_cts: np.ndarray, collect: Optional[Dict[int, np.ndarray]] = None,
                timings: Optional[dict] = None) -> np.ndarray:
    """in_cts u64 [Cin*H*W][big+1] -> output cts.  Sequence of operations identical to SURVEY A.7/A.5."""
    tlu_p, bit_p = keys.params[0], keys.params[1]
    words = in_cts.shape[-1]
    vals = {circ.input_id: in_cts.reshape(*circ.input_shape, words)}
    exact = getattr(circ, "rounding_method", "exact") == "exact"
    lsbs_after = {op.src: op.chan_lsbs() for op in circ.ops if op.kind == "tlu"} if exact else {}
    t_lin = t_ks = t_pbs = 0.0
    for op in circ.ops:
        if op.kind == "conv":
            bias = _body_constants(op.offset, lsbs_after.get(op.dst), op.acc_bits, op.out_shape[0])
            t0 = time.time()
            vals[op.dst] = O.conv2d(vals[op.src], op.weight, op.stride, op.pad, bias, depthwise=op.depthwise)
            t_lin += time.time() - t0
        elif op.kind == "lin":
            # linear combination of window taps and tensors (MaxPool2d chains): tap = depthwise one-hot convolution
            C = op.shape[0]
            consts = _body_constants(op.offset, lsbs_after.get(op.dst), op.acc_bits, C)
            t0 = time.time()
            acc = None
            for (sv, _, ky, kx), ce in zip(op.terms, op.eff_coefs()):
                if ky >= 0:
                    wk = np.zeros((C, 1, op.kernel, op.kernel), dtype=np.int32)
                    wk[:, 0, ky, kx] = ce
                    term = O.conv2d(vals[sv], wk, op.stride, op.pad, consts if acc is None else None, depthwise=True)
                    acc = term if acc is None else O.axpby(acc, 1, term, 1)
                else:
                    acc = O.axpby(acc, 1, vals[sv], ce)
            vals[op.dst] = acc
            t_lin += time.time() - t0
        elif op.kind == "fadd":
            consts = _body_constants(op.offset, lsbs_after.get(op.dst), op.acc_bits, op.shape[0])
            t0 = time.time()
            vals[op.dst] = np.stack([O.axpby(vals[op.a][c], 1, vals[op.b][c], int(op.sb[c]), int(consts[c])) for c in range(op.shape[0])])
            t_lin += time.time() - t0
        elif op.kind == "add":
            consts = _body_constants(op.offset, lsbs_after.get(op.dst), op.acc_bits, op.shape[0])
            t0 = time.time()
            vals[op.dst] = np.stack([O.axpby(vals[op.a][c], op.sa, vals[op.b][c], op.sb, int(consts[c])) for c in range(op.shape[0])])
            t_lin += time.time() - t0
        else:
            C, H, W = op.shape
            acc = np.ascontiguousarray(vals[op.src].reshape(-1, words)).copy()
            w = op.acc_bits
            # widths may differ per channel (TluOp.chan_bits): all channels share the encoding 2^(63 - w); channel c extracts
            # only its own chan_lsbs[c] low bits and its table lookup keyswitches the ciphertext scaled by 2^(w - w_c)
            w_c = op.chan_widths() if hasattr(op, "chan_widths") else np.full(C, w, dtype=np.int64)
            l_c = op.chan_lsbs() if hasattr(op, "chan_lsbs") else np.full(C, op.lsbs, dtype=np.int64)
            hw = H * W
            row_lsbs = np.repeat(l_c, hw)
            for b in range(int(l_c.max()) if exact else 0):
                rows = np.nonzero(row_lsbs > b)[0]
                sub = np.ascontiguousarray(acc[rows])
                t0 = time.time()
                small = O.keyswitch(keys.ksk[1], sub, bit_p.ksk_base_log, bit_p.ksk_level, shift=w - b, body_offset=1 << 62)
                t1 = time.time()
                c = 1 << (62 - w + b)
                lut = np.full((1, bit_p.N), (-c) & MASK64, dtype=np.uint64)
                O.pbs(keys.bsk_f[1], bit_p.bsk_base_log, small, lut, np.zeros(sub.shape[0], np.uint32), mode=1, body_const=c, out=sub,
                      big_dim=words - 1)
                acc[rows] = sub
                t2 = time.time()
                t_ks += t1 - t0; t_pbs += t2 - t1
            t0 = time.time()
            small = np.empty((acc.shape[0], tlu_p.n + 1), dtype=np.uint64)
            row_w = np.repeat(w_c, hw)
            for wv in np.unique(w_c):
                rows = np.nonzero(row_w == wv)[0]
                small[rows] = O.keyswitch(keys.ksk[0], np.ascontiguousarray(acc[rows]), tlu_p.ksk_base_log, tlu_p.ksk_level, shift=int(w - wv))
            t1 = time.time()
            luts = np.stack([lut_poly(op.tables[c], op.keep_bits, tlu_p.N, op.out_width) for c in range(C)])
            idx = np.repeat(np.arange(C, dtype=np.uint32), H * W)
            out = O.pbs(keys.bsk_f[0], tlu_p.bsk_base_log, small, luts, idx, big_dim=words - 1)
            t2 = time.time()
            t_ks += t1 - t0; t_pbs += t2 - t1
            vals[op.dst] = out.reshape(C, H, W, words)
        if collect is not None:
            collect[op.dst] = vals[op.dst]
    if timings is not None:
        timings.update({"leveled_s": t_lin, "keyswitch_s": t_ks, "pbs_s": t_pbs})
    return vals[circ.output_id].reshape(-1, words)
