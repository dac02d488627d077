"""GPU parity tests proper: every CUDA kernel, called through the C-ABI, against the CPU oracle, word for word."""
import numpy as np
import pytest
import torch

from tfx_b200.binding import KeySet, PbsParams

pytestmark = pytest.mark.gpu

TOY = [
    PbsParams(n=48, k=1, N=512, bsk_base_log=8, bsk_level=3, ksk_base_log=4, ksk_level=5, lwe_std=2.0**-30, glwe_std=2.0**-45),
    PbsParams(n=40, k=1, N=1024, bsk_base_log=12, bsk_level=2, ksk_base_log=3, ksk_level=3, lwe_std=2.0**-30, glwe_std=2.0**-48),
    PbsParams(n=32, k=1, N=2048, bsk_base_log=15, bsk_level=2, ksk_base_log=5, ksk_level=4, lwe_std=2.0**-30, glwe_std=2.0**-50),
    PbsParams(n=24, k=1, N=4096, bsk_base_log=22, bsk_level=1, ksk_base_log=4, ksk_level=4, lwe_std=2.0**-30, glwe_std=2.0**-60),
    PbsParams(n=36, k=2, N=512, bsk_base_log=9, bsk_level=2, ksk_base_log=4, ksk_level=4, lwe_std=2.0**-30, glwe_std=2.0**-45),
    PbsParams(n=28, k=2, N=1024, bsk_base_log=10, bsk_level=3, ksk_base_log=6, ksk_level=2, lwe_std=2.0**-30, glwe_std=2.0**-50),
    PbsParams(n=20, k=2, N=2048, bsk_base_log=23, bsk_level=1, ksk_base_log=4, ksk_level=4, lwe_std=2.0**-30, glwe_std=2.0**-60),
    # N = 8192: the polynomial size 7-bit lookups need (rounding_threshold_bits=7); general kernel, twiddles in global memory
    PbsParams(n=10, k=1, N=8192, bsk_base_log=20, bsk_level=1, ksk_base_log=4, ksk_level=3, lwe_std=2.0**-30, glwe_std=2.0**-60),
    PbsParams(n=8, k=1, N=8192, bsk_base_log=14, bsk_level=2, ksk_base_log=4, ksk_level=3, lwe_std=2.0**-30, glwe_std=2.0**-60),
]
IDS = [f"N{p.N}k{p.k}l{p.bsk_level}" for p in TOY]


def oracle_keys(O, p: PbsParams, seed, set_id=0):
    big = O.gen_binary_key(seed, O.ST_BIGKEY, 0, p.big_dim)
    small = O.gen_binary_key(seed, O.ST_SMALLKEY, set_id, p.n)
    ksk = O.gen_ksk(big, small, p.ksk_base_log, p.ksk_level, p.lwe_std, seed, set_id)
    bsk = O.gen_bsk(small, big, p.k, p.N, p.bsk_base_log, p.bsk_level, p.glwe_std, seed, set_id)
    return big, small, ksk, bsk


def test_fft_forward_inverse_parity(gpu_ctx, oracle):
    rng = np.random.default_rng(1)
    for N in (512, 1024, 2048, 4096, 8192):
        polys = rng.integers(-2**20, 2**20, size=(5, N)).astype(np.float64)
        polys[1] = rng.integers(-2**62, 2**62, size=N).astype(np.float64)
        d = torch.from_numpy(polys).to(gpu_ctx.device)
        f = gpu_ctx.fft_forward(d)
        f_h = f.cpu().numpy()
        for i in range(polys.shape[0]):
            ref = oracle.fft_forward(polys[i])
            assert np.array_equal(f_h[i], ref), f"forward FFT differs N={N} poly {i}"
        back = gpu_ctx.to_host_u64(gpu_ctx.fft_inverse(f))
        for i in range(polys.shape[0]):
            ref = oracle.double_to_torus(oracle.fft_inverse(f_h[i]))
            assert np.array_equal(back[i], ref), f"inverse FFT differs N={N} poly {i}"
        # round trip of small integers is exact
        assert np.array_equal(back[0].view(np.int64), polys[0].astype(np.int64))


@pytest.mark.parametrize("p", TOY, ids=IDS)
def test_keygen_parity(gpu_ctx, oracle, p):
    seed = 0x1234 + p.N
    ks = KeySet.generate(gpu_ctx, [p], seed, keep_standard_bsk=True)
    big, small, ksk, bsk = oracle_keys(oracle, p, seed)
    assert np.array_equal(ks.get_secret(-1), big)
    assert np.array_equal(ks.get_secret(0), small)
    assert np.array_equal(ks.get_ksk(0), ksk)
    assert np.array_equal(ks.get_bsk_standard(0), bsk)
    assert np.array_equal(ks.get_bsk_fourier(0), oracle.bsk_to_fourier(bsk))
    ks.close()


@pytest.mark.parametrize("p", TOY[:3], ids=IDS[:3])
def test_encrypt_phase_parity(gpu_ctx, oracle, p):
    seed, eseed = 77, 78
    ks = KeySet.generate(gpu_ctx, [p], seed)
    big = ks.get_secret(-1)
    rng = np.random.default_rng(2)
    pts = rng.integers(0, 2**64, size=37, dtype=np.uint64)
    cts = ks.encrypt(gpu_ctx.to_device_u64(pts), 2.0**-40, eseed, first_index=5)
    ref = oracle.lwe_encrypt(big, 2.0**-40, pts, eseed, first_index=5)
    assert np.array_equal(gpu_ctx.to_host_u64(cts), ref)
    ph = gpu_ctx.to_host_u64(ks.phase(cts))
    assert np.array_equal(ph, oracle.lwe_phase(big, ref))
    err = (ph - pts).view(np.int64).astype(np.float64) / 2.0**64
    assert abs(err.std() / 2.0**-40 - 1) < 0.5
    ks.close()


@pytest.mark.parametrize("p", TOY, ids=IDS)
@pytest.mark.parametrize("B", [1, 70])
def test_keyswitch_parity(gpu_ctx, oracle, p, B):
    seed = 5
    ks = KeySet.generate(gpu_ctx, [p], seed)
    ksk = ks.get_ksk(0)
    rng = np.random.default_rng(3)
    cts = rng.integers(0, 2**64, size=(B, p.big_dim + 1), dtype=np.uint64)
    for shift, off in ((0, 0), (7, 1 << 62)):
        out = ks.keyswitch(0, gpu_ctx.to_device_u64(cts), shift=shift, body_offset=off)
        ref = oracle.keyswitch(ksk, cts, p.ksk_base_log, p.ksk_level, shift=shift, body_offset=off)
        assert np.array_equal(gpu_ctx.to_host_u64(out), ref)
    ks.close()


def test_keyswitch_integer_pipe_kernel_parity(gpu_ctx, oracle, monkeypatch):
    """the IMAD keyswitch kernel (fallback when digits or accumulators do not fit the tensor-core path) against the oracle"""
    monkeypatch.setenv("TFX_KS_IMAD", "1")
    p = TOY[2]
    ks = KeySet.generate(gpu_ctx, [p], 6)
    ksk = ks.get_ksk(0)
    rng = np.random.default_rng(9)
    cts = rng.integers(0, 2**64, size=(130, p.big_dim + 1), dtype=np.uint64)
    out = ks.keyswitch(0, gpu_ctx.to_device_u64(cts), shift=3, body_offset=77)
    assert np.array_equal(gpu_ctx.to_host_u64(out), oracle.keyswitch(ksk, cts, p.ksk_base_log, p.ksk_level, shift=3, body_offset=77))
    ks.close()


@pytest.mark.parametrize("p", TOY, ids=IDS)
def test_pbs_parity(gpu_ctx, oracle, p):
    seed = 11
    ks = KeySet.generate(gpu_ctx, [p], seed)
    bsk_f = ks.get_bsk_fourier(0)
    rng = np.random.default_rng(4)
    B, T = 9, 3
    cts = rng.integers(0, 2**64, size=(B, p.n + 1), dtype=np.uint64)
    cts[0, :3] = 0                                    # exercises the ahat == 0 skip
    luts = rng.integers(0, 2**64, size=(T, p.N), dtype=np.uint64)
    idx = rng.integers(0, T, size=B).astype(np.uint32)
    d_cts, d_luts = gpu_ctx.to_device_u64(cts), gpu_ctx.to_device_u64(luts)
    d_idx = torch.from_numpy(idx.astype(np.int32)).to(gpu_ctx.device)
    out = ks.pbs(0, d_cts, d_luts, d_idx)
    ref = oracle.pbs(bsk_f, p.bsk_base_log, cts, luts, idx)
    got = gpu_ctx.to_host_u64(out)
    assert np.array_equal(got, ref)
    # fused subtract mode
    base = rng.integers(0, 2**64, size=(B, p.big_dim + 1), dtype=np.uint64)
    d_base = gpu_ctx.to_device_u64(base)
    ks.pbs(0, d_cts, d_luts, d_idx, mode=1, body_const=12345, out=d_base)
    ref2 = oracle.pbs(bsk_f, p.bsk_base_log, cts, luts, idx, mode=1, body_const=12345, out=base.copy())
    assert np.array_equal(gpu_ctx.to_host_u64(d_base), ref2)
    ks.close()


def test_pbs_functional(gpu_ctx, oracle):
    """encrypt -> keyswitch -> PBS with a random table decrypts to table[m] for every m."""
    p = PbsParams(n=64, k=1, N=1024, bsk_base_log=8, bsk_level=3, ksk_base_log=4, ksk_level=5, lwe_std=2.0**-30, glwe_std=2.0**-50)
    ks = KeySet.generate(gpu_ctx, [p], 21)
    bits = 4
    delta = 1 << (63 - bits)
    rng = np.random.default_rng(5)
    table = rng.integers(0, 2**bits, size=2**bits).astype(np.uint64)
    box = p.N >> bits
    j = np.arange(p.N)
    slot = (j + box // 2) // box
    lut = np.where(slot < 2**bits, table[slot % 2**bits] * np.uint64(delta), (np.uint64(0) - table[0] * np.uint64(delta)))
    msgs = np.arange(2**bits, dtype=np.uint64).repeat(4)
    cts = ks.encrypt(gpu_ctx.to_device_u64(msgs * np.uint64(delta)), 2.0**-45, 99)
    small = ks.keyswitch(0, cts)
    out = ks.pbs(0, small, gpu_ctx.to_device_u64(lut[None]), torch.zeros(len(msgs), dtype=torch.int32, device=gpu_ctx.device))
    ph = gpu_ctx.to_host_u64(ks.phase(out))
    dec = ((ph + np.uint64(delta // 2)) >> np.uint64(63 - bits)) & np.uint64(2**(bits + 1) - 1)
    assert np.array_equal(dec, table[msgs.astype(np.int64)])
    ks.close()


@pytest.mark.parametrize("geom", [(5, 6, 6, 7, 3, 1, 1), (4, 8, 8, 9, 3, 2, 1), (6, 5, 5, 3, 1, 1, 0), (3, 9, 7, 10, 1, 2, 0)])
def test_conv2d_parity(gpu_ctx, oracle, geom):
    Cin, H, W, Cout, ksz, stride, pad = geom
    words = 131
    rng = np.random.default_rng(6)
    x = rng.integers(0, 2**64, size=(Cin, H, W, words), dtype=np.uint64)
    w = rng.integers(-15, 16, size=(Cout, Cin, ksz, ksz)).astype(np.int32)
    bias = rng.integers(0, 2**64, size=Cout, dtype=np.uint64)
    ref = oracle.conv2d(x, w, stride, pad, bias)
    out = gpu_ctx.conv2d(gpu_ctx.to_device_u64(x), torch.from_numpy(w).to(gpu_ctx.device), stride, pad, gpu_ctx.to_device_u64(bias))
    assert np.array_equal(gpu_ctx.to_host_u64(out), ref)
    # output-channel slice (multi-GPU partition)
    out2 = gpu_ctx.conv2d(gpu_ctx.to_device_u64(x), torch.from_numpy(w).to(gpu_ctx.device), stride, pad, gpu_ctx.to_device_u64(bias),
                          oc_range=(1, Cout - 1))
    assert np.array_equal(gpu_ctx.to_host_u64(out2), ref[1:Cout - 1])


def test_depthwise_sumpool_and_axpby_parity(gpu_ctx, oracle):
    rng = np.random.default_rng(7)
    C_, H, W, words = 10, 8, 8, 77
    x = rng.integers(0, 2**64, size=(C_, H, W, words), dtype=np.uint64)
    w = np.ones((C_, 1, 7, 7), dtype=np.int32)
    ref = oracle.conv2d(x, w, 7, 0, None, depthwise=True)
    out = gpu_ctx.conv2d(gpu_ctx.to_device_u64(x), torch.from_numpy(w).to(gpu_ctx.device), 7, 0, depthwise=True)
    assert np.array_equal(gpu_ctx.to_host_u64(out), ref)
    y = rng.integers(0, 2**64, size=x.shape, dtype=np.uint64)
    ref = oracle.axpby(x, 3, y, -32, body_const=999)
    out = gpu_ctx.axpby(gpu_ctx.to_device_u64(x), 3, gpu_ctx.to_device_u64(y), -32, body_const=999)
    assert np.array_equal(gpu_ctx.to_host_u64(out), ref)
    ref = oracle.axpby(x, 1 << 5)
    out = gpu_ctx.axpby(gpu_ctx.to_device_u64(x), 1 << 5)
    assert np.array_equal(gpu_ctx.to_host_u64(out), ref)


def test_golden_kats_on_gpu(gpu_ctx):
    """the committed regression vectors (tests/golden/oracle_kats.json) reproduced by the CUDA path alone"""
    import json, os
    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_kats.json")))
    p = PbsParams(**kats["params"])
    ks = KeySet.generate(gpu_ctx, [p], kats["seed"], keep_standard_bsk=True)
    assert int(ks.get_secret(-1).sum()) == kats["big_key_weight"] and int(ks.get_secret(0).sum()) == kats["small_key_weight"]
    assert int(np.bitwise_xor.reduce(ks.get_ksk(0).reshape(-1))) == kats["ksk_xor"]
    assert int(np.bitwise_xor.reduce(ks.get_bsk_standard(0).reshape(-1))) == kats["bsk_xor"]
    cts = ks.encrypt(gpu_ctx.to_device_u64(np.array(kats["plaintexts"], dtype=np.uint64)), 2.0**-40, kats["seed"] + 1)
    assert int(np.bitwise_xor.reduce(gpu_ctx.to_host_u64(cts).reshape(-1))) == kats["cts_xor"]
    sm = ks.keyswitch(0, cts)
    assert [int(v) for v in gpu_ctx.to_host_u64(sm)[:, -1]] == kats["ks_bodies"]
    lut = (np.arange(p.N, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))[None]
    out = gpu_ctx.to_host_u64(ks.pbs(0, sm, gpu_ctx.to_device_u64(lut), torch.zeros(6, dtype=torch.int32, device=gpu_ctx.device)))
    assert [int(v) for v in out[:, -1]] == kats["pbs_bodies"]
    assert int(np.bitwise_xor.reduce(out.reshape(-1))) == kats["pbs_xor"]
    ks.close()


def test_eval_only_keyset_roundtrip(gpu_ctx, oracle):
    """server side: keys imported as bytes (no secret) give the same PBS / keyswitch results"""
    p = TOY[1]
    ks = KeySet.generate(gpu_ctx, [p], 31)
    ev = KeySet.empty(gpu_ctx, [p])
    ev.set_ksk(0, ks.get_ksk(0)); ev.set_bsk_fourier(0, ks.get_bsk_fourier(0))
    rng = np.random.default_rng(8)
    cts = gpu_ctx.to_device_u64(rng.integers(0, 2**64, size=(5, p.big_dim + 1), dtype=np.uint64))
    a = ks.keyswitch(0, cts); b = ev.keyswitch(0, cts)
    assert torch.equal(a, b)
    luts = gpu_ctx.to_device_u64(rng.integers(0, 2**64, size=(1, p.N), dtype=np.uint64))
    idx = torch.zeros(5, dtype=torch.int32, device=gpu_ctx.device)
    assert torch.equal(ks.pbs(0, a, luts, idx), ev.pbs(0, a, luts, idx))
    with pytest.raises(Exception):
        ev.encrypt(cts[:, 0].contiguous(), 2.0**-30, 1)         # no secret key on the server side
    ks.drop_secret()
    with pytest.raises(Exception):
        ks.phase(cts)
    ks.close(); ev.close()


def test_prefix_glwe_key_sets(gpu_ctx, oracle):
    """a PBS set whose GLWE key is a prefix of the big key (k*N < big_dim): keys, PBS output rows (zero-padded to the big
    dimension, both modes) equal the oracle's, and the result decrypts under the big key"""
    pa = PbsParams(n=40, k=1, N=2048, bsk_base_log=12, bsk_level=2, ksk_base_log=4, ksk_level=5, lwe_std=2.0**-30, glwe_std=2.0**-50)
    pb = PbsParams(n=32, k=2, N=512, bsk_base_log=10, bsk_level=2, ksk_base_log=4, ksk_level=5, lwe_std=2.0**-30, glwe_std=2.0**-45)
    seed = 17
    ks = KeySet.generate(gpu_ctx, [pa, pb], seed, keep_standard_bsk=True)
    assert ks.big_dim == 2048
    big = oracle.gen_binary_key(seed, oracle.ST_BIGKEY, 0, 2048)
    small_b = oracle.gen_binary_key(seed, oracle.ST_SMALLKEY, 1, pb.n)
    assert np.array_equal(ks.get_secret(-1), big) and np.array_equal(ks.get_secret(1), small_b)
    ksk_b = oracle.gen_ksk(big, small_b, pb.ksk_base_log, pb.ksk_level, pb.lwe_std, seed, 1)
    assert np.array_equal(ks.get_ksk(1), ksk_b)                                   # keyswitch key covers all 2048 mask words
    bsk_b = oracle.gen_bsk(small_b, big, pb.k, pb.N, pb.bsk_base_log, pb.bsk_level, pb.glwe_std, seed, 1)
    assert np.array_equal(ks.get_bsk_standard(1), bsk_b)
    bsk_f = oracle.bsk_to_fourier(bsk_b)
    rng = np.random.default_rng(10)
    B = 7
    cts = rng.integers(0, 2**64, size=(B, pb.n + 1), dtype=np.uint64)
    luts = rng.integers(0, 2**64, size=(2, pb.N), dtype=np.uint64)
    idx = rng.integers(0, 2, size=B).astype(np.uint32)
    d_idx = torch.from_numpy(idx.astype(np.int32)).to(gpu_ctx.device)
    out = gpu_ctx.to_host_u64(ks.pbs(1, gpu_ctx.to_device_u64(cts), gpu_ctx.to_device_u64(luts), d_idx))
    ref = oracle.pbs(bsk_f, pb.bsk_base_log, cts, luts, idx, big_dim=2048)
    assert out.shape == (B, 2049) and np.array_equal(out, ref) and not out[:, 1024:2048].any()
    base = rng.integers(0, 2**64, size=(B, 2049), dtype=np.uint64)
    d_base = gpu_ctx.to_device_u64(base)
    ks.pbs(1, gpu_ctx.to_device_u64(cts), gpu_ctx.to_device_u64(luts), d_idx, mode=1, body_const=5, out=d_base)
    ref2 = oracle.pbs(bsk_f, pb.bsk_base_log, cts, luts, idx, mode=1, body_const=5, out=base.copy(), big_dim=2048)
    assert np.array_equal(gpu_ctx.to_host_u64(d_base), ref2)
    # functional: encrypt under the big key, keyswitch to set 1, sign-PBS, decrypt under the big key
    msgs = np.array([0, 1] * 8, dtype=np.uint64)
    enc = ks.encrypt(gpu_ctx.to_device_u64((msgs << np.uint64(63)) + np.uint64(1 << 62)), 2.0**-45, 3)
    c = 1 << 60
    sign_lut = gpu_ctx.to_device_u64(np.full((1, pb.N), (-c) % 2**64, dtype=np.uint64))
    res = ks.pbs(1, ks.keyswitch(1, enc), sign_lut, torch.zeros(16, dtype=torch.int32, device=gpu_ctx.device))
    ph = gpu_ctx.to_host_u64(ks.phase(res)).view(np.int64)
    assert np.array_equal(ph > 0, msgs.astype(bool)) and np.all(np.abs(np.abs(ph) - c) < 2**50)
    ks.close()
