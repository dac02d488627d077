"""DCT preprocessing row, CPU side.

1. The numpy oracle (oracle/dct_oracle.py) against the golden vectors produced by RUNNING THE REFERENCE
   (tests/golden/dct_golden.npz, generator tests/golden/make_dct_golden.py): bit-exact, every stage and every case.
2. The oracle's OpenCV restatements against cv2 itself on random inputs (skipped if cv2 is not importable).
3. The batched torch implementation (the product, here on the CPU device; the GPU run is in test_dct_preprocess_gpu.py)
   against the oracle and the golden vectors: integer stages bit-exact; float output within FLOAT_TOL_ULPS float32 ulps
   (the block DCT is a float64 matmul whose summation order differs from numpy's; everything after it is float32).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dct-cryptonets_b200"))

from oracle import dct_oracle as DO                      # noqa: E402
from tfx_b200 import dct_preprocess as DP                # noqa: E402

GOLDEN = np.load(os.path.join(ROOT, "tests", "golden", "dct_golden.npz"))
CASES = ["c24_s16_f4", "c48_s16_f4", "c24_s8_f4", "c24_s16_f4_rect", "c24_s14_f4_big"]
FLOAT_TOL_ULPS = 2


def assert_close_ulps(got, ref, ulps=FLOAT_TOL_ULPS):
    got, ref = np.asarray(got, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    tol = ulps * np.spacing(np.maximum(np.abs(ref), np.float32(1e-3)))
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} values differ by more than {ulps} ulp; max abs diff {np.abs(got - ref).max()}"


@pytest.mark.parametrize("tag", CASES)
def test_oracle_equals_reference_outputs(tag):
    size, fs, ch = (int(v) for v in GOLDEN[tag + ".cfg"])
    for img, ref in zip(GOLDEN[tag + ".in"], GOLDEN[tag + ".out"]):
        got = DO.preprocess(img, size, fs, ch)
        assert got.dtype == np.float32 and np.array_equal(got, ref)


def test_oracle_equals_reference_stages():
    for i, img in enumerate(GOLDEN["c24_s16_f4.in"][:2]):
        st = {}
        DO.preprocess(img, 16, 4, 24, stages=st)
        assert np.array_equal(st["resized"], GOLDEN["stage.resized"][i])
        assert np.array_equal(st["crop"], GOLDEN["stage.crop"][i])
        ycc = DO.rgb_to_ycrcb_u8(st["crop"])
        assert np.array_equal(ycc, GOLDEN["stage.ycrcb"][i])
        assert np.array_equal(DO.halve_u8(ycc[..., 1]), GOLDEN["stage.chroma_half"][i])
        assert np.array_equal(st["dct_y"], GOLDEN["stage.dct_y"][i])
        assert np.array_equal(st["dct_cb"], GOLDEN["stage.dct_cb"][i])
        assert np.array_equal(st["dct_cb_upscaled"], GOLDEN["stage.dct_cb_upscaled"][i])


def test_oracle_opencv_restatements_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for h, w, dh, dw in [(32, 32, 73, 73), (32, 32, 36, 36), (48, 40, 87, 73), (96, 128, 64, 85), (33, 47, 20, 91), (32, 32, 64, 64), (224, 224, 257, 257)]:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert np.array_equal(DO.resize_linear_u8(img, dw, dh), cv2.resize(img, dsize=(dw, dh), interpolation=cv2.INTER_LINEAR))
    grid = np.stack(np.meshgrid(np.arange(0, 256, 5), np.arange(0, 256, 3), np.arange(0, 256, 7), indexing="ij"), -1).reshape(1, -1, 3).astype(np.uint8)
    for im in (rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8), grid):
        assert np.array_equal(DO.rgb_to_ycrcb_u8(im), cv2.cvtColor(cv2.cvtColor(im, cv2.COLOR_RGB2BGR), cv2.COLOR_BGR2YCrCb))
    for n in (64, 56, 30):
        c = rng.integers(0, 256, size=(n, n), dtype=np.uint8)
        assert np.array_equal(DO.halve_u8(c), cv2.resize(c, (n // 2, n // 2)))
    for (s, d) in [(8, 16), (7, 14), (28, 56)]:
        x = rng.normal(size=(s, s, 16)) * 50
        assert np.array_equal(DO.resize_linear_f64(x, d, d), cv2.resize(x, (d, d)))
        xi = rng.integers(-1024, 1024, size=(s, s, 64)).astype(np.int16)
        assert np.array_equal(DO.resize_linear_s16(xi, d, d), cv2.resize(xi, (d, d)))


def test_block_dct_is_orthonormal():
    for size in (2, 4, 8):
        T = DO.dct_matrix(size)
        assert np.allclose(T @ T.T, np.eye(size), atol=1e-14)
    flat = np.full((8, 8), 200, dtype=np.uint8)
    co = DO.block_dct(flat, 4)
    assert np.allclose(co[..., 0], (200 - 128) * 4) and np.allclose(co[..., 1:], 0, atol=1e-12)


# ---- the torch implementation ---------------------------------------------------------------------------------
def _torch_stages(pre, imgs):
    x = torch.from_numpy(imgs).to(pre.device)
    crop = pre.resize_and_crop(x)
    Y, C1, C2 = DP.rgb_to_ycrcb_u8(crop)
    return crop, Y, C1, C2


def run_torch_against_oracle(device):
    for tag in CASES:
        size, fs, ch = (int(v) for v in GOLDEN[tag + ".cfg"])
        imgs, ref = GOLDEN[tag + ".in"], GOLDEN[tag + ".out"]
        pre = DP.DctPreprocessor(size, fs, ch, "default", device=device)
        # integer stages: bit-exact
        crop, Y, C1, C2 = _torch_stages(pre, imgs)
        for i, img in enumerate(imgs):
            st = {}
            DO.preprocess(img, size, fs, ch, stages=st)
            assert np.array_equal(crop[i].cpu().numpy(), st["crop"])
            ycc = DO.rgb_to_ycrcb_u8(st["crop"])
            assert np.array_equal(torch.stack([Y[i], C1[i], C2[i]], -1).cpu().numpy().astype(np.uint8), ycc)
            assert np.array_equal(DP.halve_plane(C1[i:i + 1])[0].cpu().numpy().astype(np.uint8), DO.halve_u8(ycc[..., 1]))
        out = pre(imgs)
        assert out.dtype == torch.float32 and tuple(out.shape) == ref.shape and out.device.type == torch.device(device).type
        assert_close_ulps(out.cpu().numpy(), ref)


def test_torch_cpu_matches_reference_outputs():
    run_torch_against_oracle("cpu")


def test_torch_resize_bit_exact_random():
    rng = np.random.default_rng(11)
    for h, w, dh, dw in [(32, 32, 73, 73), (50, 70, 64, 89), (224, 224, 257, 257), (17, 9, 40, 33)]:
        imgs = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
        got = DP.resize_bilinear_u8(torch.from_numpy(imgs), dh, dw).numpy()
        for i in range(3):
            assert np.array_equal(got[i], DO.resize_linear_u8(imgs[i], dw, dh))


def test_jpeg_path_matches_its_oracle_unpinned():
    """filter_size = 8: product vs oracle restatement of libjpeg at quality 100 (UNPINNED against the reference: TurboJPEG and
    jpeg2dct are not installable here)."""
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, size=(2, 80, 80, 3), dtype=np.uint8)
    pre = DP.DctPreprocessor(8, 8, 24, "default", device="cpu")           # 8*8 = 64 crop out of a 73-pixel resize
    out = pre(imgs).numpy()
    for i in range(2):
        assert_close_ulps(out[i], DO.preprocess(imgs[i], 8, 8, 24))
    assert out.shape == (2, 24, 8, 8)
    with pytest.raises(ValueError):
        DP.DctPreprocessor(7, 8, 24, device="cpu")(imgs)                    # 56-pixel crop: chroma planes are not whole 8x8 blocks


def test_transform_loader_mirror_and_errors():
    tl = DP.TransformLoader(16)
    tf = tl.get_composed_transform_dct_img(aug=False, filter_size=4, channels=24, dct_pattern="default", device="cpu")
    img = GOLDEN["c24_s16_f4.in"][0]
    assert_close_ulps(tf(img).numpy(), GOLDEN["c24_s16_f4.out"][0])
    mgr = DP.SimpleDataManager(16, batch_size=1)                                   # same call chain as homomorphic_eval.py:110-122
    tf_np = mgr.trans_loader.get_composed_transform_dct_np(aug=False, filter_size=4, channels=24, device="cpu")
    assert torch.equal(tf_np(img), tf(img))
    with pytest.raises(NotImplementedError):
        tl.get_composed_transform_dct_img(aug=True)
    with pytest.raises(TypeError):
        tf.batch(np.zeros((1, 32, 32, 3), dtype=np.float32))
    with pytest.raises(KeyError):
        DP.DctPreprocessor(16, 4, 25, device="cpu")


def test_quantised_inputs_identical():
    """what the encrypted circuit sees: the n_bits=5 input quantisation of product and reference outputs agree"""
    ref = GOLDEN["c24_s16_f4.out"]
    out = DP.DctPreprocessor(16, 4, 24, device="cpu")(GOLDEN["c24_s16_f4.in"]).numpy()
    scale = np.abs(ref).max() / 15
    q_ref, q_out = np.clip(np.rint(ref / scale), -15, 15), np.clip(np.rint(out / scale), -15, 15)
    assert np.array_equal(q_ref, q_out)
