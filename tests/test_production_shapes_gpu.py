"""Parity at production shape: the kernel instantiations, gadget settings, batch sizes and launch
geometry the headline circuit actually runs — pbs_kernel_v8<10,2,1> (bit extraction, base 2^23, one level) and
pbs_kernel_v8<11,1,2> (table lookups, base 2^15, two levels), the tensor-core keyswitch at (base 2^2, 5 / 8 levels,
2048-bit big key) — against the CPU oracle, word for word.  Keys are generated on the GPU and handed to the oracle
(keygen parity is tests/test_kernels_gpu.py::test_keygen_parity)."""
import os

import numpy as np
import pytest
import torch

from tfx_b200.binding import KeySet, PbsParams
from tfx_b200 import params as P

pytestmark = pytest.mark.gpu

# the sets pick_parameters returns for DCT-ResNet-20 / 24x16^2 / n_bits 5 / rounding 6 / p_error 0.01
TLU = PbsParams(n=768, k=1, N=2048, bsk_base_log=15, bsk_level=2, ksk_base_log=2, ksk_level=8,
                lwe_std=P.min_noise_std(768), glwe_std=P.min_noise_std(2048))
BIT = PbsParams(n=492, k=2, N=1024, bsk_base_log=23, bsk_level=1, ksk_base_log=2, ksk_level=5,
                lwe_std=P.min_noise_std(492), glwe_std=P.min_noise_std(2048))


def _short(p: PbsParams, n: int) -> PbsParams:
    """same kernel instantiation and gadget, fewer CMux steps (keeps the CPU side cheap for big batches)"""
    return PbsParams(n=n, k=p.k, N=p.N, bsk_base_log=p.bsk_base_log, bsk_level=p.bsk_level, ksk_base_log=p.ksk_base_log,
                     ksk_level=p.ksk_level, lwe_std=p.lwe_std, glwe_std=p.glwe_std)


def _pbs_case(gpu_ctx, oracle, p, B, seed, tables=3):
    ks = KeySet.generate(gpu_ctx, [p], seed)
    bsk_f = ks.get_bsk_fourier(0)
    rng = np.random.default_rng(seed)
    cts = rng.integers(0, 2**64, size=(B, p.n + 1), dtype=np.uint64)
    cts[0, :2] = 0                                                    # the ahat == 0 skip
    luts = rng.integers(0, 2**64, size=(tables, p.N), dtype=np.uint64)
    idx = rng.integers(0, tables, size=B).astype(np.uint32)
    d_cts, d_luts = gpu_ctx.to_device_u64(cts), gpu_ctx.to_device_u64(luts)
    d_idx = torch.from_numpy(idx.astype(np.int32)).to(gpu_ctx.device)
    got = gpu_ctx.to_host_u64(ks.pbs(0, d_cts, d_luts, d_idx))
    ref = oracle.pbs(bsk_f, p.bsk_base_log, cts, luts, idx)
    bad = np.flatnonzero((got != ref).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {B} ciphertexts differ (first rows {bad[:8].tolist()})"
    # accumulate mode (the bit-extraction step subtracts the PBS result from the accumulator ciphertext in place)
    base = rng.integers(0, 2**64, size=(B, p.big_dim + 1), dtype=np.uint64)
    d_base = gpu_ctx.to_device_u64(base)
    ks.pbs(0, d_cts, d_luts, d_idx, mode=1, body_const=1 << 50, out=d_base)
    ref2 = oracle.pbs(bsk_f, p.bsk_base_log, cts, luts, idx, mode=1, body_const=1 << 50, out=base.copy())
    assert np.array_equal(gpu_ctx.to_host_u64(d_base), ref2)
    ks.close()


@pytest.mark.parametrize("p,n_short", [(BIT, 12), (TLU, 10)], ids=["bit_k2_N1024_l1_b23", "tlu_k1_N2048_l2_b15"])
def test_pbs_batch_three_times_the_resident_grid(gpu_ctx, oracle, p, n_short):
    """B > 3x the resident grid (4 or 2 CTAs on each of 148 SMs): every CTA takes several ciphertexts from the dynamic
    hand-out loop and re-initialises its accumulator in between."""
    resident = 148 * (4 if p.N <= 1024 else 2)
    _pbs_case(gpu_ctx, oracle, _short(p, n_short), 3 * resident + 37, seed=31)


@pytest.mark.parametrize("p", [BIT, TLU], ids=["bit", "tlu"])
def test_pbs_hand_out_loop_with_a_capped_grid(gpu_ctx, oracle, p, monkeypatch):
    """grid capped to 4 CTAs (TFX_PBS_GRID_CAP): 60 ciphertexts -> 15 per CTA, mid-sized n"""
    monkeypatch.setenv("TFX_PBS_GRID_CAP", "4")
    _pbs_case(gpu_ctx, oracle, _short(p, 40), 60, seed=32)


@pytest.mark.parametrize("p", [BIT, TLU], ids=["bit", "tlu"])
def test_pbs_general_kernel_same_shapes(gpu_ctx, oracle, p, monkeypatch):
    """the general kernel (pbs_kernel, any gadget) stays available behind TFX_PBS_V7=1 and must agree too"""
    monkeypatch.setenv("TFX_PBS_V7", "1")
    _pbs_case(gpu_ctx, oracle, _short(p, 16), 70, seed=33)


@pytest.fixture(scope="module")
def shipped(gpu_ctx):
    ks = KeySet.generate(gpu_ctx, [TLU, BIT], 41)
    yield ks
    ks.close()


@pytest.mark.parametrize("sid", [0, 1], ids=["tlu_set", "bit_set"])
def test_shipped_sets_pbs_and_keyswitch_300(gpu_ctx, oracle, shipped, sid):
    """exact shipped sets (n = 768 / 492), B = 300: crosses the tensor-core keyswitch's 128-row tiles and runs full-length blind rotations"""
    ks, p = shipped, (TLU, BIT)[sid]
    rng = np.random.default_rng(50 + sid)
    B = 300
    big = rng.integers(0, 2**64, size=(B, ks.big_dim + 1), dtype=np.uint64)
    ksk = ks.get_ksk(sid)
    for shift, off in ((0, 0), (7, 1 << 62)):
        got = gpu_ctx.to_host_u64(ks.keyswitch(sid, gpu_ctx.to_device_u64(big), shift=shift, body_offset=off))
        ref = oracle.keyswitch(ksk, big, p.ksk_base_log, p.ksk_level, shift=shift, body_offset=off)
        assert np.array_equal(got, ref), f"keyswitch differs (shift {shift})"
    small = ref
    bsk_f = ks.get_bsk_fourier(sid)
    luts = rng.integers(0, 2**64, size=(2, p.N), dtype=np.uint64)
    idx = rng.integers(0, 2, size=B).astype(np.uint32)
    out = ks.pbs(sid, gpu_ctx.to_device_u64(small), gpu_ctx.to_device_u64(luts), torch.from_numpy(idx.astype(np.int32)).to(gpu_ctx.device))
    ref = oracle.pbs(bsk_f, p.bsk_base_log, small, luts, idx, big_dim=ks.big_dim)
    assert np.array_equal(gpu_ctx.to_host_u64(out), ref)


def test_rounding_chain_and_lookup_of_a_1536_row_slice(gpu_ctx, oracle, shipped):
    """One rank's share of a headline layer at 8 GPUs: 1536 accumulator ciphertexts (2.6 waves of the bit kernel) through
    two exact bit extractions (keyswitch with shift and offset, sign PBS subtracted in place) and, for the first 512 rows,
    the table lookup — the executor's chain (executor.py) with the kernels on one side and the oracle on the other."""
    from tfx_b200.executor import bit_lut, lut_polynomials
    ks = shipped
    rows, w, keep, lsbs, rows_tlu = 1536, 8, 6, 2, 512
    rng = np.random.default_rng(60)
    # fresh encryptions of w-bit accumulator values at delta = 2^(63 - w), plus the half-LSB rounding offset
    acc_vals = rng.integers(0, (1 << w) - 4, size=rows).astype(np.uint64)     # (+ half LSB stays below the padding bit)
    pts = ((acc_vals + np.uint64(1 << (lsbs - 1))) << np.uint64(63 - w))
    d_acc = ks.encrypt(gpu_ctx.to_device_u64(pts), TLU.glwe_std, 61)
    acc = gpu_ctx.to_host_u64(d_acc).copy()
    ksk_bit, ksk_tlu = ks.get_ksk(1), ks.get_ksk(0)
    bsk_bit, bsk_tlu = ks.get_bsk_fourier(1), ks.get_bsk_fourier(0)
    zero = torch.zeros(rows, dtype=torch.int32, device=gpu_ctx.device)
    for b in range(lsbs):
        lut, c = bit_lut(w, b, BIT.N)
        d_small = ks.keyswitch(1, d_acc, shift=w - b, body_offset=1 << 62)
        ks.pbs(1, d_small, gpu_ctx.to_device_u64(lut[None]), zero, mode=1, body_const=c, out=d_acc)
        small = oracle.keyswitch(ksk_bit, acc, BIT.ksk_base_log, BIT.ksk_level, shift=w - b, body_offset=1 << 62)
        assert np.array_equal(gpu_ctx.to_host_u64(d_small), small), f"bit step {b}: keyswitch differs"
        acc = oracle.pbs(bsk_bit, BIT.bsk_base_log, small, lut[None], np.zeros(rows, np.uint32), mode=1, body_const=c, out=acc,
                         big_dim=ks.big_dim)
        assert np.array_equal(gpu_ctx.to_host_u64(d_acc), acc), f"bit step {b}: accumulators differ"
    tables = rng.integers(-16, 16, size=(4, 1 << keep)).astype(np.int64)
    luts = lut_polynomials(tables, keep, TLU.N, 6)
    idx = (np.arange(rows_tlu) % 4).astype(np.uint32)
    # the low bits are gone: the 6 kept bits already sit right under the padding bit (2^(63-w) * 2^lsbs = 2^(63-keep)), no shift
    d_small = ks.keyswitch(0, d_acc[:rows_tlu].contiguous())
    small = oracle.keyswitch(ksk_tlu, acc[:rows_tlu], TLU.ksk_base_log, TLU.ksk_level)
    assert np.array_equal(gpu_ctx.to_host_u64(d_small), small)
    out = ks.pbs(0, d_small, gpu_ctx.to_device_u64(luts), torch.from_numpy(idx.astype(np.int32)).to(gpu_ctx.device))
    ref = oracle.pbs(bsk_tlu, TLU.bsk_base_log, small, luts, idx, big_dim=ks.big_dim)
    assert np.array_equal(gpu_ctx.to_host_u64(out), ref)
    # and the chain did what it is for: the decrypted lookup equals the table of the exactly rounded accumulator
    ph = gpu_ctx.to_host_u64(ks.phase(out))
    dec = ((ph + (np.uint64(1) << np.uint64(56))) >> np.uint64(57)).astype(np.int64) & 127
    dec = np.where(dec >= 64, dec - 128, dec)
    rounded = ((acc_vals[:rows_tlu].astype(np.int64) + (1 << (lsbs - 1))) >> lsbs) & ((1 << keep) - 1)
    expect = tables[idx, rounded]
    assert (dec == expect).mean() > 0.97                              # p_error = 0.01 per PBS by design


@pytest.mark.parametrize("env", ["TFX_KS_IMMA", "TFX_KS_IMAD"], ids=["mma_sync_fallback", "integer_pipe_fallback"])
def test_keyswitch_fallback_kernels_at_the_shipped_sets(gpu_ctx, oracle, shipped, monkeypatch, env):
    """the tcgen05 keyswitch is the default; the mma.sync kernel (TFX_KS_IMMA=1) and the integer-pipe kernel (TFX_KS_IMAD=1) stay
    available for gadgets / devices the tensor path cannot hold and must give the same words"""
    monkeypatch.setenv(env, "1")
    ks = shipped
    rng = np.random.default_rng(70)
    big = rng.integers(0, 2**64, size=(150, ks.big_dim + 1), dtype=np.uint64)
    for sid, p in enumerate((TLU, BIT)):
        got = gpu_ctx.to_host_u64(ks.keyswitch(sid, gpu_ctx.to_device_u64(big), shift=5, body_offset=1 << 62))
        ref = oracle.keyswitch(ks.get_ksk(sid), big, p.ksk_base_log, p.ksk_level, shift=5, body_offset=1 << 62)
        assert np.array_equal(got, ref)

