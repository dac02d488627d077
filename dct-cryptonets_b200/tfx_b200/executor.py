"""GPU executor of a Circuit: keygen, encrypt, run (leveled ops + rounding chains + table lookups), decrypt.

The layer loop is host Python; every arithmetic step is a call into libtfx_b200.so on the context's stream
(binding.py).  Replaces what Circuit.keygen / encrypt_run_decrypt do behind reference homomorphic_eval.py:315,:70.

Multi-GPU: keys are replicated (every rank generates the same keys from the same seed, no
broadcast needed), each table-lookup layer is partitioned by output channel over the ranks — a rank computes the
leveled op, the rounding chain and the PBS only for its own channels — and the layer output is re-assembled with
an all-gather over NCCL.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .binding import Context, KeySet, PbsParams, TfxError, launch_count
from .circuit import Circuit, ConvOp, AddOp, LinOp, TluOp, channel_offsets

MASK64 = (1 << 64) - 1
TLU_SET, BIT_SET = 0, 1


def lut_polynomials(tables: np.ndarray, keep_bits: int, N: int, out_width: int) -> np.ndarray:
    """tables int64 [C][2^p] -> torus LUT polynomials u64 [C][N]: every entry repeated N/2^p times,
    rotated by half a box; the top half-box holds -T[0] (negacyclic wrap).  Output delta = 2^(63 - out_width)."""
    C, size = tables.shape
    assert size == 1 << keep_bits and N % size == 0
    box = N // size
    j = np.arange(N)
    slot = (j + box // 2) // box                      # 0 .. 2^p
    delta = np.uint64(1) << np.uint64(63 - out_width)
    vals = tables.astype(np.int64).view(np.uint64) * delta          # two's complement wrap is the torus encoding
    lut = vals[:, slot % size]
    wrap = slot >= size
    lut[:, wrap] = (np.uint64(0) - vals[:, 0:1]).repeat(int(wrap.sum()), axis=1)
    return np.ascontiguousarray(lut)


def bit_lut(acc_bits: int, b: int, N: int) -> Tuple[np.ndarray, int]:
    """constant LUT -c with c = 2^(62 - w + b): PBS gives -c / +c, adding c gives bit_b * 2^(63 - w + b)"""
    c = 1 << (62 - acc_bits + b)
    return np.full(N, (-c) & MASK64, dtype=np.uint64), c


def channel_range(C: int, rank: int, world: int) -> Tuple[int, int, int]:
    """output-channel block [lo, hi) owned by `rank` and the padded block size (SURVEY §8e partition)"""
    per = (C + world - 1) // world
    lo = min(C, rank * per)
    hi = min(C, lo + per)
    return lo, hi, per


def gather_channels(local: torch.Tensor, C: int, per: int, hw: int, world: int, pg=None) -> torch.Tensor:
    """local [(hi-lo)*hw][words] -> full [C*hw][words] on every rank (all-gather of equal, zero-padded blocks)"""
    if world == 1:
        return local
    import torch.distributed as dist
    words = local.shape[-1]
    pad_rows = per * hw
    if local.shape[0] != pad_rows:
        buf = torch.zeros(pad_rows, words, dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
        local = buf
    full = torch.empty(world * pad_rows, words, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, local.contiguous(), group=pg)
    return full[: C * hw]


@dataclass
class RunStats:
    seconds: float = 0.0
    pbs_tlu: int = 0
    pbs_bit: int = 0
    keyswitches: int = 0
    launches: int = 0
    layer_seconds: Optional[List[Tuple[str, float]]] = None
    kernel_events: Optional[list] = None          # (class, event0, event1, units) per launch when profiling

    def kernel_seconds(self) -> Dict[str, Tuple[float, int, int]]:
        """class -> (seconds, launches, units); CUDA-event time on the launching stream, call after a synchronize"""
        out: Dict[str, Tuple[float, int, int]] = {}
        for cls, e0, e1, units in (self.kernel_events or []):
            s, n, u = out.get(cls, (0.0, 0, 0))
            out[cls] = (s + e0.elapsed_time(e1) / 1e3, n + 1, u + units)
        return out


class CircuitExecutor:
    def __init__(self, circuit: Circuit, params: Tuple[PbsParams, PbsParams], ctx: Optional[Context] = None,
                 rank: int = 0, world_size: int = 1, process_group=None, input_std: Optional[float] = None):
        self.circ = circuit
        self.params = list(params)
        if ctx is None:
            if not torch.cuda.is_available():
                raise TfxError("CUDA device required: the TFHE execution path has no CPU fallback "
                               "(fhe='disable' / 'simulate' run without a GPU; keygen() and fhe='execute' do not)")
            ctx = Context(torch.cuda.current_device())
        self.ctx = ctx
        self.rank, self.world = rank, world_size
        self.pg = process_group
        self.big_dim = max(p.big_dim for p in params)
        self.words = self.big_dim + 1
        self.input_std = input_std if input_std is not None else params[0].glwe_std
        self.keys: Optional[KeySet] = None
        self._luts: Dict[int, torch.Tensor] = {}
        self._lut_index: Dict[int, torch.Tensor] = {}
        self._bit_luts: Dict[Tuple[int, int], Tuple[torch.Tensor, int]] = {}
        self._weights: Dict[int, torch.Tensor] = {}
        self._bias: Dict[int, torch.Tensor] = {}
        # Wave-sized chains on several streams: with W ranks a layer of 12 288 ciphertexts leaves 12 288 / W per kernel — 2.6
        # waves of the bit-extraction PBS at W = 8 — and every one of the ~14 launches of a rounding chain would pay its own
        # partial last wave.  A rank's share of a layer is therefore cut into chunks of whole waves (WAVE_ROWS ciphertexts = the
        # resident grid of the bit kernel = two waves of the table kernel) whose chains are enqueued on separate streams: when
        # one chunk's kernel runs out of ciphertexts the next step of another chunk is already queued and takes the freed SMs,
        # so only the end of the layer sees idle SMs.  On by default for world_size > 1; TFX_SPLIT_STREAMS=0 switches it off,
        # TFX_SPLIT_STREAMS=k (k >= 2) allows up to k concurrent chains.  Results are identical (the chunks are independent
        # ciphertexts).
        env = os.environ.get("TFX_SPLIT_STREAMS")
        self.max_chains = (4 if world_size > 1 else 1) if env is None else max(1, int(env) if env != "1" else 2)
        self.split_streams = self.max_chains > 1
        self._sides: List[Tuple[torch.cuda.Stream, Context]] = []
        if self.ctx.device.type == "cuda":               # one wave of the bit-extraction kernel: 4 resident CTAs per SM
            self.WAVE_ROWS = 4 * torch.cuda.get_device_properties(self.ctx.device).multi_processor_count
        self._perm: Dict[Tuple[int, int, int], Tuple[torch.Tensor, np.ndarray]] = {}
        self._prepare_constants()

    # ---- constants ------------------------------------------------------------------------------------------
    def _prepare_constants(self):
        dev = self.ctx.device
        N_tlu, N_bit = self.params[TLU_SET].N, self.params[BIT_SET].N
        for op in self.circ.ops:
            if op.kind == "conv":
                self._weights[op.dst] = torch.from_numpy(np.ascontiguousarray(op.weight)).to(dev)
                self._bias[op.dst] = self.ctx.to_device_u64(self._body_constants(op, op.out_shape[0]))
            elif op.kind == "lin":
                # one depthwise one-hot kernel per window tap, carrying the term's (shifted) coefficient
                C = op.shape[0]
                for idx, ((sv, _, ky, kx), ce) in enumerate(zip(op.terms, op.eff_coefs())):
                    if ky >= 0:
                        wk = np.zeros((C, 1, op.kernel, op.kernel), dtype=np.int32)
                        wk[:, 0, ky, kx] = ce
                        self._weights[(op.dst, idx)] = torch.from_numpy(wk).to(dev)
                self._bias[op.dst] = self.ctx.to_device_u64(self._body_constants(op, C))
            elif op.kind == "tlu":
                luts = lut_polynomials(op.tables, op.keep_bits, N_tlu, op.out_width)
                self._luts[op.dst] = self.ctx.to_device_u64(luts)
                C, H, W = op.shape
                idx = np.repeat(np.arange(C, dtype=np.int32), H * W)
                self._lut_index[op.dst] = torch.from_numpy(idx).to(dev)
                for b in range(op.lsbs):
                    key = (op.acc_bits, b)
                    if key not in self._bit_luts:
                        lut, c = bit_lut(op.acc_bits, b, N_bit)
                        self._bit_luts[key] = (self.ctx.to_device_u64(lut[None]), c)
        self._zero_idx = torch.zeros(max(int(np.prod(op.shape)) for op in self.circ.lookups()), dtype=torch.int32, device=dev)

    def _body_constants(self, lin_op, channels: int) -> np.ndarray:
        """u64 [C]: (offset_c + half LSB of the rounding that follows, per channel) at the accumulator's encoding, added to the body word"""
        ls = self._lsbs_after(lin_op, channels)
        offs = channel_offsets(lin_op.offset, channels)
        return np.array([((int(o) + ((1 << (int(l) - 1)) if l > 0 else 0)) << (63 - lin_op.acc_bits)) & MASK64 for o, l in zip(offs, ls)],
                        dtype=np.uint64)

    def _lsbs_after(self, lin_op, channels: int) -> np.ndarray:
        """int64 [C]: rounding bits removed, per channel, by the lookup that consumes this accumulator (0 if it is the circuit
        output).  Decides the half-LSB offset folded into the accumulator: in approximate mode the LUT's own half-box rotation
        already rounds to nearest, so no offset is added."""
        if self.circ.rounding_method == "exact":
            for op in self.circ.ops:
                if op.kind == "tlu" and op.src == lin_op.dst:
                    return op.chan_lsbs()
        return np.zeros(channels, dtype=np.int64)

    # ---- keys ------------------------------------------------------------------------------------------------
    def _fresh_seed(self) -> bytes:
        """16 bytes from the OS CSPRNG; with several ranks rank 0 draws and every rank receives the same bytes (the keys
        and the input ciphertexts are replicated)"""
        seed = os.urandom(16)
        if self.world > 1:
            import torch.distributed as dist
            box = [seed]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
            seed = box[0]
        return seed

    def _bind_stream(self):
        """library kernels and the torch ops around them (gathers, permutations, NCCL) must share one stream: follow
        torch's current stream on every entry point instead of the one that was current at construction"""
        if self.ctx.device.type == "cuda":                 # (tests drive the host logic with a CPU stand-in context)
            self.ctx.set_stream(torch.cuda.current_stream(self.ctx.device))

    def keygen(self, seed=None, keep_standard_bsk: bool = False) -> float:
        """seed: 16 bytes or an int; None (the default) draws a fresh one from the OS CSPRNG like Concrete's keygen.
        A fixed seed is for tests and benchmarks only: the secret keys are a deterministic function of it."""
        t0 = time.time()
        self._bind_stream()
        if self.keys is not None:
            self.keys.close()
        if seed is None:
            seed = self._fresh_seed()
        self.keys = KeySet.generate(self.ctx, self.params, seed, keep_standard_bsk)
        self._enc_seed, self._enc_next = None, 0
        self.ctx.synchronize()
        return time.time() - t0

    def use_keys(self, keys: KeySet):
        self.keys = keys
        self._enc_seed, self._enc_next = None, 0

    # ---- client side -------------------------------------------------------------------------------------------
    def encrypt(self, q_in: np.ndarray, enc_seed=None, first_index: Optional[int] = None) -> torch.Tensor:
        """q_in int64 [C][H][W] -> ciphertexts int64-view u64 [C*H*W][big_dim+1] at the input encoding width.
        enc_seed None (the default): masks and noise come from a per-key-set random seed whose PRF index advances by the
        number of ciphertexts already produced, so no two ciphertexts ever share a mask.  An explicit enc_seed (tests,
        benchmarks) reproduces the same ciphertexts on every call."""
        self._bind_stream()
        delta_shift = 63 - self.circ.input_width
        pts = (q_in.astype(np.int64).reshape(-1).view(np.uint64) << np.uint64(delta_shift))
        d_pts = self.ctx.to_device_u64(pts)
        if enc_seed is None:
            if getattr(self, "_enc_seed", None) is None:
                self._enc_seed, self._enc_next = self._fresh_seed(), 0
            enc_seed, first_index = self._enc_seed, self._enc_next
            self._enc_next += int(pts.size)
        return self.keys.encrypt(d_pts, self.input_std, enc_seed, first_index or 0)

    def decrypt(self, cts: torch.Tensor) -> np.ndarray:
        self._bind_stream()
        ph = self.ctx.to_host_u64(self.keys.phase(cts))
        w = self.circ.output_width
        shift = np.uint64(63 - w)
        u = ((ph + (np.uint64(1) << (shift - np.uint64(1)))) >> shift) & np.uint64((1 << (w + 1)) - 1)
        u = u.astype(np.int64)
        if self.circ.output_is_acc:
            offs = channel_offsets(self.circ.output_offset, self.circ.output_shape[0] if len(self.circ.output_shape) > 1 else 1)
            return u - (np.repeat(offs, u.size // offs.size) if offs.size > 1 else int(offs[0]))
        return np.where(u >= (1 << w), u - (1 << (w + 1)), u)      # signed two's complement in w+1 bits

    # ---- server side ---------------------------------------------------------------------------------------------
    def _channel_range(self, C: int) -> Tuple[int, int, int]:
        return channel_range(C, self.rank, self.world)

    def _sorted_rows(self, dst: int, lo: int, hi: int, w_c: np.ndarray, hw: int):
        """(device permutation sorted row -> original row, channel order) for a lookup layer with per-channel widths"""
        key = (dst, lo, hi)
        if key not in self._perm:
            order = np.argsort(-w_c, kind="stable")
            perm = (order[:, None] * hw + np.arange(hw)[None, :]).reshape(-1).astype(np.int64)
            self._perm[key] = (torch.from_numpy(perm).to(self.ctx.device), order)
        return self._perm[key]

    WAVE_ROWS = 592          # default: 148 SMs x 4 resident CTAs of the bit-extraction PBS kernel (= 2 waves of the table kernel)

    def _side(self, i: int) -> Tuple[torch.cuda.Stream, Context]:
        while len(self._sides) <= i:
            st = torch.cuda.Stream(self.ctx.device)
            with torch.cuda.stream(st):
                self._sides.append((st, Context(self.ctx.device.index)))   # binds to the side stream; own scratch + hand-out counter
        return self._sides[i]

    def _chunks(self, nloc: int) -> Tuple[List[Tuple[int, int]], Optional[Tuple[int, int]]]:
        """(row ranges of the concurrent chains, row range of the chain that runs ahead or None).
        The chains are whole waves of the bit-extraction kernel (WAVE_ROWS) each, at most max_chains of them.  What is left over a
        whole number of table-kernel waves (WAVE_ROWS / 2 ciphertexts; the table lookup is the LAST launch of a layer, 7 ms per
        ciphertext) becomes a small chain of its own that is enqueued completely before the others: its lookup runs early, beside
        the other chains' extraction steps, and the lookups that end the layer fill the GPU exactly instead of leaving a partial
        last wave to drain before the all-gather."""
        if self.max_chains <= 1 or nloc <= self.WAVE_ROWS:
            return [(0, nloc)], None
        mode = os.environ.get("TFX_CHUNK_MODE", "ahead")
        if mode == "equal":
            per = -(-nloc // self.max_chains)
            return [(r0, min(nloc, r0 + per)) for r0 in range(0, nloc, per)], None
        ahead = None
        body = nloc
        if mode == "ahead":
            rem = nloc % (self.WAVE_ROWS // 2)
            if 0 < rem < nloc:
                ahead, body = (nloc - rem, nloc), nloc - rem          # the narrowest channels (rows are sorted widest first)
        waves = -(-body // self.WAVE_ROWS)
        per = -(-waves // max(1, self.max_chains - (1 if ahead else 0))) * self.WAVE_ROWS
        return [(r0, min(body, r0 + per)) for r0 in range(0, body, per)], ahead

    def _gather(self, local: torch.Tensor, C: int, per: int, hw: int) -> torch.Tensor:
        return gather_channels(local, C, per, hw, self.world, self.pg)

    def run(self, in_cts: torch.Tensor, stats: Optional[RunStats] = None, time_layers: bool = False,
            profile_kernels: bool = False) -> torch.Tensor:
        """in_cts [Cin*H*W][words] -> output ciphertexts [n_out][words].  Enqueues on the context stream."""
        circ, ctx, keys = self.circ, self.ctx, self.keys
        assert keys is not None, "keygen() first"
        self._bind_stream()
        words = self.words
        prof = profile_kernels and stats is not None
        if prof and stats.kernel_events is None:
            stats.kernel_events = []

        def timed(cls, units, fn):
            if not prof:
                return fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(); e1.record()
            stats.kernel_events.append((cls, e0, e1, units))
            return r

        vals: Dict[int, torch.Tensor] = {circ.input_id: in_cts.view(*circ.input_shape, words)}
        acc_local: Dict[int, Tuple[torch.Tensor, int, int, int]] = {}     # acc id -> (local acc, lo, hi, per)
        launches0 = launch_count()
        layer_t = []
        ev0 = None
        # liveness: a layer tensor (hundreds of MB to GB) is released after its last consumer
        last_use: Dict[int, int] = {}
        for i, op in enumerate(circ.ops):
            for src in ((op.a, op.b) if op.kind in ("add", "fadd") else [t_[0] for t_ in op.terms] if op.kind == "lin" else (op.src,)):
                last_use[src] = i
        last_use[circ.output_id] = len(circ.ops)
        for i, op in enumerate(circ.ops):
            if time_layers:
                ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
            if op.kind == "conv":
                C = op.out_shape[0]
                is_out = (op.dst == circ.output_id)
                lo, hi, per = (0, C, C) if is_out else self._channel_range(C)
                if hi > lo:
                    acc = timed("conv", (hi - lo) * op.out_shape[1] * op.out_shape[2],
                                lambda: ctx.conv2d(vals[op.src], self._weights[op.dst], op.stride, op.pad, self._bias[op.dst],
                                                   oc_range=(lo, hi), depthwise=op.depthwise))
                else:
                    acc = ctx.empty_u64(0, op.out_shape[1], op.out_shape[2], words)
                acc_local[op.dst] = (acc.view(-1, words), lo, hi, per)
                if is_out:
                    vals[op.dst] = acc
            elif op.kind == "add":
                C, H, W = op.shape
                lo, hi, per = self._channel_range(C)
                consts = self._body_constants(op, C)
                if hi > lo:
                    a = vals[op.a][lo:hi].contiguous()
                    b = vals[op.b][lo:hi].contiguous()
                    if len(set(int(v) for v in consts[lo:hi])) == 1:
                        acc = timed("add", a.shape[0] * H * W, lambda: ctx.axpby(a, op.sa, b, op.sb, body_const=int(consts[lo])))
                    else:                                            # per-channel offsets: one launch per channel (H*W ciphertexts each)
                        acc = torch.empty_like(a)
                        for c in range(hi - lo):
                            timed("add", H * W, lambda: ctx.axpby(a[c], op.sa, b[c], op.sb, body_const=int(consts[lo + c]), out=acc[c]))
                else:
                    acc = ctx.empty_u64(0, H, W, words)
                acc_local[op.dst] = (acc.view(-1, words), lo, hi, per)
            elif op.kind == "lin":
                # leveled linear combination of window taps and tensors (MaxPool2d chains, circuit.LinOp): taps are depthwise one-hot
                # convolutions carrying the coefficient (the first one also adds the per-channel body constants), the rest axpby
                C, H, W = op.shape
                lo, hi, per = self._channel_range(C)
                if hi > lo:
                    acc = None
                    for idx, ((sv, _, ky, kx), ce) in enumerate(zip(op.terms, op.eff_coefs())):
                        if ky >= 0:
                            term = timed("conv", (hi - lo) * H * W,
                                         lambda: ctx.conv2d(vals[sv], self._weights[(op.dst, idx)], op.stride, op.pad,
                                                            self._bias[op.dst] if acc is None else None, oc_range=(lo, hi), depthwise=True))
                            acc = term if acc is None else timed("add", (hi - lo) * H * W, lambda: ctx.axpby(acc, 1, term, 1))
                        else:
                            term = vals[sv][lo:hi].contiguous()
                            acc = timed("add", (hi - lo) * H * W, lambda: ctx.axpby(acc, 1, term, ce))
                else:
                    acc = ctx.empty_u64(0, H, W, words)
                acc_local[op.dst] = (acc.view(-1, words), lo, hi, per)
            elif op.kind == "fadd":
                # fused residual add (opt-in, circuit.FusedAddOp): this rank's conv accumulators + m_c * shortcut, per channel
                conv_acc, lo, hi, per = acc_local.pop(op.a)
                C, H, W = op.shape
                consts = self._body_constants(op, C)
                if hi > lo:
                    a = conv_acc.view(hi - lo, H, W, words)
                    b = vals[op.b][lo:hi].contiguous()
                    acc = torch.empty_like(a)
                    for c in range(hi - lo):
                        timed("add", H * W, lambda: ctx.axpby(a[c], 1, b[c], int(op.sb[lo + c]), body_const=int(consts[lo + c]), out=acc[c]))
                else:
                    acc = ctx.empty_u64(0, H, W, words)
                acc_local[op.dst] = (acc.view(-1, words), lo, hi, per)
            else:
                acc, lo, hi, per = acc_local.pop(op.src)
                C, H, W = op.shape
                nloc = acc.shape[0]
                if nloc > 0:
                    w = op.acc_bits
                    hw = H * W
                    exact = circ.rounding_method == "exact"
                    lut_idx = self._lut_index[op.dst][lo * hw: hi * hw]
                    w_c = op.chan_widths()[lo:hi]
                    l_c = op.chan_lsbs()[lo:hi] if exact else np.zeros(hi - lo, dtype=np.int64)
                    uniform = bool((w_c == w).all())
                    if uniform:
                        src, out, idx_rows = acc, ctx.empty_u64(nloc, words), lut_idx
                        steps = [nloc] * (op.lsbs if exact else 0)                  # rows [0, steps[b]) take part in extraction step b
                        segments = [(0, nloc, w)]                                  # (row range, width) for the table lookup's keyswitch
                    else:
                        # per-channel widths: every channel removes only its own low bits.  Rows are sorted by channel width
                        # (widest first), so the rows still active in step b are a prefix and equal-width rows are contiguous.
                        perm_d, order = self._sorted_rows(op.dst, lo, hi, w_c, hw)
                        src = acc.index_select(0, perm_d)
                        idx_rows = lut_idx.index_select(0, perm_d)
                        out = ctx.empty_u64(nloc, words)
                        l_s, w_s = l_c[order], w_c[order]
                        steps = [hw * int((l_s > b).sum()) for b in range(int(l_s.max()))]
                        segments, c0 = [], 0
                        for c1 in range(1, len(w_s) + 1):
                            if c1 == len(w_s) or w_s[c1] != w_s[c0]:
                                segments.append((c0 * hw, c1 * hw, int(w_s[c0])))
                                c0 = c1
                    n_small = self.params[TLU_SET].n + 1

                    def bit_step(c_, r0, r1, b):
                        """extraction step b of rows [r0, r1) of this rank's share (only the rows still active), on c_'s stream"""
                        e_ = min(r1, steps[b])
                        if e_ <= r0:
                            return
                        a_, n_ = src[r0:e_], e_ - r0
                        small = timed("ks_bit", n_, lambda: keys.keyswitch(BIT_SET, a_, shift=w - b, body_offset=1 << 62, ctx=c_))
                        lut, c = self._bit_luts[(w, b)]
                        timed("pbs_bit", n_, lambda: keys.pbs(BIT_SET, small, lut, self._zero_idx[:n_], mode=1, body_const=c, out=a_, ctx=c_))

                    def lookup_step(c_, r0, r1):
                        small = c_.empty_u64(r1 - r0, n_small)
                        for s0, s1, wv in segments:                                # one keyswitch per width: the ciphertext is scaled by 2^(w - wv)
                            g0, g1 = max(s0, r0), min(s1, r1)
                            if g1 > g0:
                                timed("ks_tlu", g1 - g0, lambda: keys.keyswitch(TLU_SET, src[g0:g1], shift=w - wv, out=small[g0 - r0: g1 - r0], ctx=c_))
                        timed("pbs_tlu", r1 - r0, lambda: keys.pbs(TLU_SET, small, self._luts[op.dst], idx_rows[r0:r1], out=out[r0:r1], ctx=c_))

                    chunks, ahead = self._chunks(nloc)
                    if len(chunks) > 1 or ahead is not None:
                        # one stream per chunk; the steps are enqueued round-robin over the chunks so that the chains advance
                        # together (the GPU serves older grids first): a chunk's partial last wave is filled by the next chunk's
                        # CTAs of the same step, and the first chunk's next step is queued by the time the last chunk drains
                        main = torch.cuda.current_stream(ctx.device)
                        lanes = [(main, ctx)] + [self._side(i) for i in range(len(chunks) - 1 + (1 if ahead else 0))]
                        for st, _ in lanes[1:]:
                            st.wait_stream(main)                     # acc / out are produced / allocated on the main stream
                        if ahead is not None:                        # the leftover chain runs ahead, complete, on the last lane
                            st, sctx = lanes[-1]
                            with torch.cuda.stream(st):
                                for b in range(len(steps)):
                                    bit_step(sctx, ahead[0], ahead[1], b)
                                lookup_step(sctx, ahead[0], ahead[1])
                        for b in range(len(steps)):
                            for (st, sctx), (r0, r1) in zip(lanes, chunks):
                                with torch.cuda.stream(st):
                                    bit_step(sctx, r0, r1, b)
                        for (st, sctx), (r0, r1) in zip(lanes, chunks):
                            with torch.cuda.stream(st):
                                lookup_step(sctx, r0, r1)
                        for st, _ in lanes[1:]:
                            main.wait_stream(st)
                    else:
                        for b in range(len(steps)):
                            bit_step(ctx, 0, nloc, b)
                        lookup_step(ctx, 0, nloc)
                    if not uniform:                                  # back to channel order: out[perm[i]] = sorted_out[i]
                        unsorted = torch.empty_like(out)
                        unsorted.index_copy_(0, perm_d, out)
                        out = unsorted
                    if stats is not None:
                        stats.pbs_bit += hw * int(l_c.sum()); stats.pbs_tlu += nloc; stats.keyswitches += hw * int(l_c.sum()) + nloc
                else:
                    out = ctx.empty_u64(0, words)
                del acc
                # layer exchange (NCCL all-gather on the compute stream); timed as class 'gather' when profiling
                vals[op.dst] = (timed("gather", nloc, lambda: self._gather(out, C, per, H * W)) if self.world > 1
                                else self._gather(out, C, per, H * W)).view(C, H, W, words)
            if time_layers:
                ev1 = torch.cuda.Event(enable_timing=True); ev1.record()
                layer_t.append((op.name, ev0, ev1))
            for vid in [v for v in vals if last_use.get(v, -1) <= i and v != circ.output_id]:
                del vals[vid]
        out = vals[circ.output_id].reshape(-1, words)
        if stats is not None:
            stats.launches += launch_count() - launches0
            if time_layers:
                torch.cuda.synchronize()
                stats.layer_seconds = [(n, a.elapsed_time(b) / 1e3) for n, a, b in layer_t]
        return out
