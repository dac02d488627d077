"""The oracle against mathematics (it has no reference vectors to be pinned to — parity unpinned):
exact negacyclic products, decomposition bounds, decrypt round trips, noise statistics against the model, every
message through keyswitch + PBS, the exact rounding chain against integer rounding, and committed regression KATs."""
import json
import math
import os

import numpy as np
import pytest

from tfx_b200 import params as P
from tfx_b200.binding import PbsParams

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_kats.json")


def test_decomposition_bounds_and_recomposition(oracle):
    rng = np.random.default_rng(0)
    xs = rng.integers(0, 2**64, size=4000, dtype=np.uint64)
    xs[:4] = [0, 2**64 - 1, 2**63, 2**63 - 1]
    for bl, lv in ((2, 7), (4, 5), (8, 3), (16, 2), (24, 1), (12, 3), (1, 10)):
        d = oracle.decompose(xs, bl, lv)
        B = 1 << bl
        assert d.min() >= -B // 2 and d.max() < B // 2 or (d.max() <= B // 2)
        assert d.min() >= -(B // 2) and d.max() <= B // 2 - 1
        rec = np.zeros(xs.size, dtype=np.uint64)
        for j in range(lv):
            rec += d[:, j].astype(np.uint64) << np.uint64(64 - bl * (j + 1))
        err = (xs - rec).view(np.int64)
        assert np.abs(err).max() <= 1 << (63 - bl * lv)


def test_negacyclic_fft_product_is_exact_for_small_operands(oracle):
    rng = np.random.default_rng(1)
    for N in (512, 1024, 4096):
        a = rng.integers(-2**9, 2**9, size=N)
        b = rng.integers(-2**20, 2**20, size=N)
        fa, fb = oracle.fft_forward(a.astype(np.float64)), oracle.fft_forward(b.astype(np.float64))
        ca, cb = fa[:, 0] + 1j * fa[:, 1], fb[:, 0] + 1j * fb[:, 1]
        prod = ca * cb
        got = oracle.double_to_torus(oracle.fft_inverse(np.stack([prod.real, prod.imag], 1)))
        want = oracle.poly_mul_negacyclic(a.astype(np.int64).view(np.uint64), b.astype(np.int64).view(np.uint64))
        assert np.array_equal(got, want)
        back = oracle.fft_inverse(fa)
        assert np.abs(back - a).max() < 2**-30


def test_fft_rounding_noise_law(oracle):
    """external-product FFT error: sigma^2 ~ 2^-3.3 * 2^22 * l*(k+1) * N * B^2 / 2^128 (params.var_fft_extprod uses 2^-3.0)"""
    rng = np.random.default_rng(2)
    N, bl = 1024, 16
    errs = []
    for _ in range(4):
        d = rng.integers(-2**(bl - 1), 2**(bl - 1), size=N)
        key = rng.integers(0, 2**64, size=N, dtype=np.uint64)
        fd, fk = oracle.fft_forward(d.astype(np.float64)), oracle.fft_forward(key.view(np.int64).astype(np.float64))
        prod = (fd[:, 0] + 1j * fd[:, 1]) * (fk[:, 0] + 1j * fk[:, 1])
        got = oracle.double_to_torus(oracle.fft_inverse(np.stack([prod.real, prod.imag], 1)))
        want = oracle.poly_mul_negacyclic(d.astype(np.int64).view(np.uint64), key)
        errs.append((got - want).view(np.int64).astype(np.float64) / 2.0**64)
    var = np.concatenate(errs).var()
    model = P.var_fft_extprod(0, N, bl, 1)          # one product: level * (k+1) = 1
    assert var < model and var > model / 16


def test_gaussian_and_prf(oracle):
    g = oracle.gauss_fill(3, 9, 0, 100000)
    assert abs(g.mean()) < 0.02 and abs(g.std() - 1) < 0.02 and abs(((g - g.mean())**4).mean() / g.var()**2 - 3) < 0.15
    u = oracle.prf_fill(3, 9, 0, 4096)
    assert np.array_equal(u[100:200], oracle.prf_fill(3, 9, 100, 100))
    assert not np.array_equal(u[:100], oracle.prf_fill(3, 10, 0, 100))
    bits = np.unpackbits(u.view(np.uint8))
    assert abs(bits.mean() - 0.5) < 0.01


def _toy():
    return PbsParams(n=64, k=1, N=1024, bsk_base_log=8, bsk_level=3, ksk_base_log=4, ksk_level=5, lwe_std=2.0**-24, glwe_std=2.0**-40)


@pytest.fixture(scope="module")
def toy_keys(oracle):
    p = _toy()
    big = oracle.gen_binary_key(7, oracle.ST_BIGKEY, 0, p.big_dim)
    small = oracle.gen_binary_key(7, oracle.ST_SMALLKEY, 0, p.n)
    ksk = oracle.gen_ksk(big, small, p.ksk_base_log, p.ksk_level, p.lwe_std, 7)
    bsk = oracle.bsk_to_fourier(oracle.gen_bsk(small, big, p.k, p.N, p.bsk_base_log, p.bsk_level, p.glwe_std, 7))
    return p, big, small, ksk, bsk


def test_encrypt_decrypt_and_noise_std(oracle, toy_keys):
    p, big, *_ = toy_keys
    rng = np.random.default_rng(3)
    pts = rng.integers(0, 2**64, size=2000, dtype=np.uint64)
    cts = oracle.lwe_encrypt(big, 2.0**-30, pts, 11)
    err = (oracle.lwe_phase(big, cts) - pts).view(np.int64).astype(np.float64) / 2.0**64
    assert abs(err.std() / 2.0**-30 - 1) < 0.1 and abs(err.mean()) < 2.0**-32
    assert 0.4 < big.mean() < 0.6


def test_keyswitch_noise_matches_model(oracle, toy_keys):
    p, big, small, ksk, _ = toy_keys
    rng = np.random.default_rng(4)
    pts = rng.integers(0, 2**64, size=1500, dtype=np.uint64)
    cts = oracle.lwe_encrypt(big, 2.0**-50, pts, 12)
    out = oracle.keyswitch(ksk, cts, p.ksk_base_log, p.ksk_level)
    err = (oracle.lwe_phase(small, out) - pts).view(np.int64).astype(np.float64) / 2.0**64
    model = math.sqrt(P.var_keyswitch(p))
    assert 0.7 < err.std() / model < 1.3


def test_pbs_every_message_and_output_noise(oracle, toy_keys):
    p, big, small, ksk, bsk = toy_keys
    rng = np.random.default_rng(5)
    for bits in (1, 2, 3, 4, 5):
        delta = 1 << (63 - bits)
        table = rng.integers(0, 2**bits, size=2**bits).astype(np.uint64)
        box = p.N >> bits
        slot = (np.arange(p.N) + box // 2) // box
        lut = np.where(slot < 2**bits, table[slot % 2**bits] * np.uint64(delta),
                       (np.zeros(1, np.uint64) - table[0] * np.uint64(delta))[0]).astype(np.uint64)
        msgs = np.arange(2**bits, dtype=np.uint64).repeat(max(1, 64 >> bits))
        cts = oracle.lwe_encrypt(big, 2.0**-45, msgs * np.uint64(delta), 13 + bits)
        out = oracle.pbs(bsk, p.bsk_base_log, oracle.keyswitch(ksk, cts, p.ksk_base_log, p.ksk_level), lut[None],
                         np.zeros(msgs.size, np.uint32))
        ph = oracle.lwe_phase(big, out)
        want = table[msgs.astype(np.int64)] * np.uint64(delta)
        dec = ((ph + np.uint64(delta // 2)) >> np.uint64(63 - bits)) & np.uint64(2**(bits + 1) - 1)
        assert np.array_equal(dec, table[msgs.astype(np.int64)]), f"bits={bits}"
        if bits == 3:
            err = (ph - want).view(np.int64).astype(np.float64) / 2.0**64
            assert err.std() < 1.5 * math.sqrt(P.var_pbs_out(p))


def test_exact_rounding_chain_equals_integer_rounding(oracle, toy_keys):
    """SURVEY A.7 on every value of a 9-bit accumulator rounded to 6 bits, then an identity table."""
    p, big, small, ksk, bsk = toy_keys
    w, t = 9, 6
    lsbs = w - t
    u = np.arange(0, 2**w - (1 << (lsbs - 1)), dtype=np.uint64)           # values whose rounding stays inside t bits
    delta_w = np.uint64(1) << np.uint64(63 - w)
    half = np.uint64(1 << (lsbs - 1))
    acc = oracle.lwe_encrypt(big, 2.0**-50, (u + half) * delta_w, 21)
    for b in range(lsbs):
        smallct = oracle.keyswitch(ksk, acc, p.ksk_base_log, p.ksk_level, shift=w - b, body_offset=1 << 62)
        c = 1 << (62 - w + b)
        lut = np.full((1, p.N), (-c) % 2**64, dtype=np.uint64)
        oracle.pbs(bsk, p.bsk_base_log, smallct, lut, np.zeros(u.size, np.uint32), mode=1, body_const=c, out=acc)
    ph = oracle.lwe_phase(big, acc)
    delta_t = 1 << (63 - t)
    dec = ((ph + np.uint64(delta_t // 2)) >> np.uint64(63 - t)) & np.uint64(2**(t + 1) - 1)
    want = (u + half) >> np.uint64(lsbs)
    assert np.array_equal(dec, want)


def test_leveled_conv_on_ciphertexts_equals_integer_conv(oracle, toy_keys):
    p, big, *_ = toy_keys
    rng = np.random.default_rng(6)
    Cin, H, W, Cout = 3, 5, 5, 4
    x = rng.integers(0, 16, size=(Cin, H, W))
    wts = rng.integers(-7, 8, size=(Cout, Cin, 3, 3)).astype(np.int32)
    bits = 12
    delta = np.uint64(1) << np.uint64(63 - bits)
    cts = oracle.lwe_encrypt(big, 2.0**-50, x.reshape(-1).astype(np.uint64) * delta, 31).reshape(Cin, H, W, -1)
    out = oracle.conv2d(cts, wts, 1, 1)
    ph = oracle.lwe_phase(big, out.reshape(-1, out.shape[-1]))
    dec = (((ph + (delta >> np.uint64(1))) >> np.uint64(63 - bits)) & np.uint64(2**(bits + 1) - 1)).astype(np.int64)
    dec = np.where(dec >= 2**bits, dec - 2**(bits + 1), dec).reshape(Cout, H, W)
    import torch, torch.nn.functional as F
    want = F.conv2d(torch.from_numpy(x[None].astype(np.float64)), torch.from_numpy(wts.astype(np.float64)), padding=1)[0].numpy().astype(np.int64)
    assert np.array_equal(dec, want)


def test_regression_kats(oracle):
    """Committed known-answer vectors produced by tests/golden/make_oracle_kats.py from this oracle; they pin the
    conventions the CUDA kernels were verified against (they are NOT reference vectors)."""
    kats = json.load(open(GOLDEN))
    assert [int(v) for v in oracle.prf_fill(kats["seed"], 5, 3, 4)] == kats["prf"]
    assert [int(v) for v in oracle.decompose(np.array(kats["decompose_in"], dtype=np.uint64), 7, 3).reshape(-1)] == kats["decompose_out"]
    p = PbsParams(**kats["params"])
    big = oracle.gen_binary_key(kats["seed"], oracle.ST_BIGKEY, 0, p.big_dim)
    small = oracle.gen_binary_key(kats["seed"], oracle.ST_SMALLKEY, 0, p.n)
    assert int(big.sum()) == kats["big_key_weight"] and int(small.sum()) == kats["small_key_weight"]
    ksk = oracle.gen_ksk(big, small, p.ksk_base_log, p.ksk_level, p.lwe_std, kats["seed"])
    bsk = oracle.gen_bsk(small, big, p.k, p.N, p.bsk_base_log, p.bsk_level, p.glwe_std, kats["seed"])
    assert int(np.bitwise_xor.reduce(ksk.reshape(-1))) == kats["ksk_xor"]
    assert int(np.bitwise_xor.reduce(bsk.reshape(-1))) == kats["bsk_xor"]
    cts = oracle.lwe_encrypt(big, 2.0**-40, np.array(kats["plaintexts"], dtype=np.uint64), kats["seed"] + 1)
    assert int(np.bitwise_xor.reduce(cts.reshape(-1))) == kats["cts_xor"]
    sm = oracle.keyswitch(ksk, cts, p.ksk_base_log, p.ksk_level)
    assert [int(v) for v in sm[:, -1]] == kats["ks_bodies"]
    lut = (np.arange(p.N, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))[None]
    out = oracle.pbs(oracle.bsk_to_fourier(bsk), p.bsk_base_log, sm, lut, np.zeros(sm.shape[0], np.uint32))
    assert [int(v) for v in out[:, -1]] == kats["pbs_bodies"]
    assert int(np.bitwise_xor.reduce(out.reshape(-1))) == kats["pbs_xor"]
