"""CPU tests of the host logic: compile front-end, clear evaluator, parameter picker, LUT construction (against the
oracle's own builder), wire format, and the reference-facing API surface."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from tfx_b200 import circuit as C
from tfx_b200 import params as P
from tfx_b200.binding import PbsParams
from tfx_b200.resnet_dct import ResidualBlock, resnet20_dct, resnet18_dct


class TinyNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.trunk = nn.Sequential(nn.Conv2d(3, 4, 1, bias=False), nn.BatchNorm2d(4), nn.ReLU(), ResidualBlock(4, 4, False),
                                   ResidualBlock(4, 6, True), nn.AvgPool2d(2), nn.Flatten())
        self.final_feat_dim = 6

    def forward(self, x):
        return self.trunk(x)


@pytest.fixture(scope="module")
def tiny():
    torch.manual_seed(0)
    m = TinyNet().eval()
    calib = torch.randn(64, 3, 4, 4)
    return m, calib, C.build_circuit(m, calib, 5, 6, 0.01)


def test_tiny_structure(tiny):
    m, calib, circ = tiny
    kinds = [op.kind for op in circ.ops]
    assert kinds.count("conv") == 7 and kinds.count("add") == 2 and kinds.count("tlu") == 8   # 6 convs + sum-pool
    assert circ.output_is_acc and circ.output_shape == (6,)
    for op in circ.lookups():
        assert op.keep_bits == min(op.acc_bits, 6) and op.tables.shape == (op.shape[0], 1 << op.keep_bits)
        assert op.out_width >= 1
    text = circ.to_text()
    assert "table_lookup" in text and "conv2d" in text and text.count("\n") == len(circ.ops) + 1


def test_clear_evaluation_tracks_float_model(tiny):
    m, calib, circ = tiny
    x = calib[:16]
    y = C.dequantize_output(circ, C.evaluate_clear(circ, C.quantize_input(circ, x.numpy())))
    ref = m(x).detach().numpy()
    assert y.shape == ref.shape
    assert np.corrcoef(y.ravel(), ref.ravel())[0, 1] > 0.9


def test_accumulators_fit_their_width_on_calibration(tiny):
    m, calib, circ = tiny
    vals = {}
    q = C.quantize_input(circ, calib.numpy())
    C.evaluate_clear(circ, q, collect=vals)
    for op in circ.ops:
        if op.kind in ("conv", "add"):
            u = vals[op.dst] + C.channel_offsets(op.offset, vals[op.dst].shape[1]).reshape(1, -1, 1, 1)
            assert u.min() >= 0 and u.max() < (1 << op.acc_bits)


def test_per_channel_offsets_are_centred_and_robust():
    """One offset per channel, one width per tensor: every channel is centred in [0, 2^w) (equal slack on both sides — a channel
    whose minimum sits at 0 wraps below zero at the first noisy or out-of-calibration value and returns a negated table entry),
    fewer bit extractions than a tensor-wide offset, and the noise simulation on unseen images stays free of such wraps."""
    torch.manual_seed(0)
    model = resnet20_dct(24, 16).eval()
    g = torch.Generator().manual_seed(0)
    calib = torch.randn(40, 24, 16, 16, generator=g)
    unseen = torch.randn(3, 24, 16, 16, generator=g).numpy()
    per = C.build_circuit(model, calib, 5, 6, 0.01, per_channel_widths=False)
    wide = C.build_circuit(model, calib, 5, 6, 0.01, per_channel_offsets=False)
    both = C.build_circuit(model, calib, 5, 6, 0.01)                                  # default: per-channel offsets and widths
    assert per.pbs_count()["tlu"] == wide.pbs_count()["tlu"] == both.pbs_count()["tlu"]
    assert both.pbs_count()["bit"] < 0.95 * per.pbs_count()["bit"] < 0.95 * 0.97 * wide.pbs_count()["bit"]
    assert all(op.chan_bits is None for op in per.ops if op.kind != "tlu" or True) and any(op.chan_bits is not None for op in both.lookups())
    vals = {}
    C.evaluate_clear(per, C.quantize_input(per, calib.numpy()), collect=vals)
    for op in per.ops:
        if op.kind in ("conv", "add"):
            u = vals[op.dst] + C.channel_offsets(op.offset, vals[op.dst].shape[1]).reshape(1, -1, 1, 1)
            lo, hi = u.min(axis=(0, 2, 3)), ((1 << op.acc_bits) - 1) - u.max(axis=(0, 2, 3))
            assert (lo >= 0).all() and (hi >= 0).all()
            lsbs = max(0, op.acc_bits - 6)
            assert (np.abs(lo - hi) <= (1 << lsbs) // 2 + 2).all()          # centred up to the rounding half and integer division
    for circ in (per, both):
        tlu, bit, _ = P.pick_parameters(circ.noise_spec())
        nm = C.NoiseModel.from_params(tlu, bit, tlu.glwe_std)
        q = C.quantize_input(circ, unseen)
        clear = C.evaluate_clear(circ, q)
        span = int(clear.max() - clear.min())
        for seed in range(2):
            noisy = C.evaluate_clear(circ, q, noise=nm, rng=np.random.default_rng(seed))
            assert np.abs(noisy - clear).max() < 0.2 * span
    # per-channel widths: every channel fits its own width with symmetric slack
    vals = {}
    C.evaluate_clear(both, C.quantize_input(both, calib.numpy()), collect=vals)
    for op in both.lookups():
        lin = next(o for o in both.ops if getattr(o, "dst", None) == op.src)
        u = vals[op.src] + C.channel_offsets(lin.offset, vals[op.src].shape[1]).reshape(1, -1, 1, 1)
        top = (1 << op.chan_widths()) - 1
        assert (u.min(axis=(0, 2, 3)) >= 0).all() and (u.max(axis=(0, 2, 3)) <= top).all()
        assert (op.chan_widths() <= op.acc_bits).all() and (op.chan_widths() >= min(op.acc_bits, 6)).all()


def test_rounding_semantics_of_tlu_apply():
    # a 9-bit accumulator rounded to 6 bits: index = round_half_up(u / 8); padding-bit wrap negates
    tables = np.arange(64, dtype=np.int64).reshape(1, 64) * 3 + 1
    op = C.TluOp("t", 0, 1, (1, 1, 1), 9, 6, tables, C.QuantInfo(1.0, 0, 255))
    acc = np.arange(-20, 492, dtype=np.int64).reshape(-1, 1, 1, 1)
    off = 20
    got = C.tlu_apply(op, off, acc).reshape(-1)
    u = np.arange(0, 512)
    idx = (u + 4) >> 3
    want = np.where(idx >= 64, -(tables[0][idx % 64]), tables[0][idx % 64])
    assert np.array_equal(got, want)


def test_lut_polynomials_match_oracle_builder(oracle):
    from oracle import circuit_oracle as CO
    from tfx_b200.executor import lut_polynomials, bit_lut
    rng = np.random.default_rng(0)
    for keep, N, width in ((6, 4096, 12), (4, 512, 7), (6, 2048, 13), (3, 1024, 4)):
        tables = rng.integers(-31, 32, size=(3, 1 << keep)).astype(np.int64)
        got = lut_polynomials(tables, keep, N, width)
        for c in range(3):
            assert np.array_equal(got[c], CO.lut_poly(tables[c], keep, N, width))
    lut, c = bit_lut(12, 3, 2048)
    assert c == 1 << 53 and np.all(lut == np.uint64((-c) % 2**64))


def test_resnet20_workload_numbers():
    torch.manual_seed(0)
    circ = C.build_circuit(resnet20_dct(24, 16).eval(), torch.randn(8, 24, 16, 16), 5, 6, 0.01)
    # 21 convs, 89.2 M MACs (+ the 64x49 sum-pool), 307 200 table-lookup elements when the stem quantiser folds
    assert sum(1 for op in circ.ops if op.kind == "conv" and not op.depthwise) == 21
    assert circ.macs() == 89_243_648 + 64 * 49
    assert circ.pbs_count()["tlu"] == 307_200
    assert circ.maximum_integer_bit_width() <= 16


def test_resnet18_added_stem_key_compiles():
    torch.manual_seed(0)
    circ = C.build_circuit(resnet18_dct(24, 16).eval(), torch.randn(4, 24, 16, 16), 5, 6, 0.01)
    assert circ.output_shape == (512,)


def test_parameter_picker_meets_noise_constraints(tiny):
    m, calib, circ = tiny
    spec = circ.noise_spec()
    tlu, bit, info = P.pick_parameters(spec)
    assert tlu.k * tlu.N == info["big_dim"] and bit.k * bit.N <= tlu.k * tlu.N      # the bit set may use a prefix of the big key
    # the big key dimension is searched: forcing the largest one is feasible too but never cheaper
    tlu4, bit4, _ = P.pick_parameters(spec, big_dim=4096)
    assert tlu4.k * tlu4.N == 4096 and P._cost(spec, tlu4, bit4) >= P._cost(spec, tlu, bit)
    ok, margin = P._check(spec, tlu, bit, P.z_score(spec.p_error))
    assert ok and margin >= 1.0 and abs(info["z"] - 2.5758) < 1e-3
    # a smaller p_error must never yield cheaper parameters; an unreachable one must fail loudly
    spec2 = circ.noise_spec(); spec2.p_error = 1e-3
    tlu2, bit2, _ = P.pick_parameters(spec2)
    assert P._cost(spec2, tlu2, bit2) >= P._cost(spec, tlu, bit) and P._check(spec2, tlu2, bit2, P.z_score(1e-3))[0]
    # (N = 8192 made p_error = 1e-30 reachable for this small circuit; 11-bit lookups are beyond every supported polynomial size)
    spec3 = circ.noise_spec(); spec3.p_error = 1e-30
    for lk in spec3.lookups:
        lk.acc_bits, lk.keep_bits = max(lk.acc_bits, 12), 11
    with pytest.raises(ValueError):
        P.pick_parameters(spec3)


def test_seven_bit_lookups_get_the_8192_polynomial(tiny):
    """rounding_threshold_bits=7 (the reference's ImageNet setting, run_homomorphic_eval.sh:25): the mod-switch noise at 2N = 8192
    alone exceeds a 7-bit lookup's budget, so the picker must move the table set to (k=1, N=8192) — served by the general PBS kernel"""
    m, calib, _ = tiny
    circ7 = C.build_circuit(m, calib, 5, 7, 0.01)
    assert max(op.keep_bits for op in circ7.lookups()) == 7
    tlu, bit, info = P.pick_parameters(circ7.noise_spec())
    assert (tlu.k, tlu.N) == (1, 8192) and info["big_dim"] == 8192 and bit.k * bit.N <= 8192
    assert P._check(circ7.noise_spec(), tlu, bit, P.z_score(0.01))[0]
    with pytest.raises(ValueError):
        P.pick_parameters(circ7.noise_spec(), big_dim=4096)


def test_work_formulas_match_survey_examples():
    p = PbsParams(860, 1, 4096, 22, 1, 3, 3, 0.0, 0.0)
    assert abs(P.pbs_flops(p) / 0.44e9 - 1) < 0.02 and abs(P.bsk_bytes(p) / 113e6 - 1) < 0.01   # SURVEY §8(d) worked example
    assert P.ks_macs(p) == 4096 * 3 * 861


def test_concrete_api_surface(tiny):
    from concrete.fhe import Configuration
    from concrete.ml.torch.compile import compile_brevitas_qat_model, compile_torch_model
    m, calib, _ = tiny
    cfg = Configuration(show_progress=False, progress_tag=True, progress_title="Evaluation: ")
    qm = compile_torch_model(m, calib, rounding_threshold_bits=6, p_error=0.01, n_bits=5, configuration=cfg, verbose=False)
    assert qm.fhe_circuit.graph.maximum_integer_bit_width() <= 16
    assert isinstance(qm.fhe_circuit.mlir, str) and len(qm.fhe_circuit.mlir) > 100
    x = calib[:3].numpy()
    y = qm.forward(x, fhe="simulate")
    y0 = qm.forward(x, fhe="disable")
    assert y.shape == (3, 6) and np.abs(y - y0).max() <= 0.35 * np.abs(y0).max()      # noisy simulation tracks the clear values
    with pytest.raises(ValueError):
        qm.forward(x, fhe="bogus")
    qa = compile_torch_model(m, calib, rounding_threshold_bits={"n_bits": 6, "method": "approximate"}, p_error=0.01, n_bits=5)
    assert qa.fhe_circuit.statistics["bit"] == 0 and qa.fhe_circuit.statistics["tlu"] == qm.fhe_circuit.statistics["tlu"]
    assert "approximate" in qa.fhe_circuit.mlir
    with pytest.raises(ValueError):
        compile_torch_model(m, calib, rounding_threshold_bits={"n_bits": 6, "method": "nearest"}, p_error=0.01, n_bits=5)
    qm2 = compile_brevitas_qat_model(m, calib, rounding_threshold_bits=6, n_bits=5, p_error=0.01, configuration=cfg)
    assert qm2.fhe_circuit.statistics["total"] == qm.fhe_circuit.statistics["total"]
    if not torch.cuda.is_available():
        with pytest.raises(Exception):          # fhe='execute' must fail loudly without the CUDA path (no CPU fallback)
            qm.forward(x[:1], fhe="execute")


def test_wire_format_roundtrip():
    from concrete.ml.deployment.fhe_client_server import _pack, _unpack
    a = np.arange(12, dtype=np.uint64).reshape(3, 4)
    b = np.linspace(0, 1, 5)
    h, (a2, b2) = _unpack(_pack({"kind": "x", "v": 3}, [a, b]))
    assert h["kind"] == "x" and np.array_equal(a, a2) and np.array_equal(b, b2)
    with pytest.raises(ValueError):
        _unpack(b"nope" + bytes(20))


def test_circuit_bundle_is_pickle_free_and_lossless(tiny, tmp_path):
    """client.zip / server.zip carry the circuit as a JSON header + raw arrays: loading one executes no code (a bundle comes
    from the model provider and is opened by the process that holds the secret key), and the circuit survives unchanged."""
    import zipfile
    from types import SimpleNamespace
    from concrete.ml.deployment.fhe_client_server import FHEModelDev, _load_bundle, _pack
    m, calib, circ = tiny
    fused = C.build_circuit(m, calib, 5, 6, 0.01, fuse_residual=True)
    tlu, bit, _ = P.pick_parameters(circ.noise_spec())
    x = C.quantize_input(circ, calib[:4].numpy())
    for i, cc in enumerate((circ, fused)):
        d = tmp_path / f"bundle{i}"
        FHEModelDev(str(d), SimpleNamespace(fhe_circuit=SimpleNamespace(circuit=cc, params=(tlu, bit)))).save()
        with zipfile.ZipFile(d / "client.zip") as z:
            assert not any(n.endswith(".pkl") for n in z.namelist())
        for name in ("client.zip", "server.zip"):
            c2, params = _load_bundle(str(d), name)
            assert list(params) == [tlu, bit]
            assert c2.to_text() == cc.to_text()
            assert np.array_equal(C.evaluate_clear(c2, x), C.evaluate_clear(cc, x))
    head, arrays = C.circuit_to_portable(circ)
    head["ops"][0]["evil"] = 1
    with pytest.raises(ValueError):
        C.circuit_from_portable(head, arrays)


REF = "/root/reference/dct-cryptonets"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_topology_matches_reference_modules():
    """Our ResNet-20 restatement and the reference's ResNetDCT (models/backbone.py:107-184,291-302) give the same
    outputs for the same weights, hence compile to the same circuit."""
    stub = types.ModuleType("brevitas"); nn_stub = types.ModuleType("brevitas.nn"); q_stub = types.ModuleType("brevitas.quant")
    for name in ("QuantConv2d", "QuantReLU", "QuantIdentity"):
        setattr(nn_stub, name, type(name, (nn.Module,), {}))
    q_stub.Int8ActPerTensorFloat = q_stub.Int8WeightPerTensorFloat = object
    stub.nn, stub.quant = nn_stub, q_stub
    saved = {k: sys.modules.get(k) for k in ("brevitas", "brevitas.nn", "brevitas.quant", "models", "models.backbone")}
    sys.modules.update({"brevitas": stub, "brevitas.nn": nn_stub, "brevitas.quant": q_stub})
    sys.path.insert(0, REF)
    try:
        sys.modules.pop("models", None); sys.modules.pop("models.backbone", None)
        from models import backbone
        torch.manual_seed(3)
        ref = backbone.ResNet20(in_channels=24, img_size=16).eval()
        ours = resnet20_dct(24, 16).eval()
        missing = ours.load_state_dict(ref.state_dict(), strict=True)
        x = torch.randn(2, 24, 16, 16)
        assert torch.allclose(ours(x), ref(x), atol=1e-5)
        assert ours.final_feat_dim == ref.final_feat_dim
        # the RGB stem with a max-pool (backbone.py:447-455: 7x7 stride-2 conv, ReLU, MaxPool2d(3, 2, 1)), ResNet-18
        ref18 = backbone.ResNet18(in_channels=3, img_size=128).eval()
        ours18 = resnet18_dct(3, 128).eval()
        ours18.load_state_dict(ref18.state_dict(), strict=True)
        x = torch.randn(1, 3, 128, 128)
        assert torch.allclose(ours18(x), ref18(x), atol=1e-5)
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_noisy_simulation_error_rate_matches_p_error(tiny):
    """fhe='simulate' draws the modelled noise: with negligible variances it equals the clear evaluation; with the picked
    parameters a single table lookup fails about as often as p_error allows (not orders of magnitude more or less)."""
    m, calib, circ = tiny
    q = C.quantize_input(circ, calib.numpy())
    clear = C.evaluate_clear(circ, q)
    quiet = C.NoiseModel(1e-40, 1e-40, 1e-40, 1e-40, 1e-40)
    assert np.array_equal(C.evaluate_clear(circ, q, noise=quiet), clear)
    tlu, bit, _ = P.pick_parameters(circ.noise_spec())
    nm = C.NoiseModel.from_params(tlu, bit, tlu.glwe_std)
    op = circ.lookups()[1]
    lin = next(o for o in circ.ops if getattr(o, "dst", None) == op.src)
    lk = circ.noise_spec().lookups[1]
    vals = {}
    C.evaluate_clear(circ, q, collect=vals)
    acc = np.repeat(vals[op.src], 8, axis=0)
    want = C.tlu_apply(op, lin.offset, acc)
    got = C.tlu_apply_noisy(op, lin.offset, acc, True, lk.weight_norm2, lk.fresh_inputs, nm, np.random.default_rng(1))
    idx_changed = (got != want).mean()
    assert idx_changed < 0.2            # at most ~ (lsbs + 1) * p_error, and many flips do not change the table value
    noisy = C.evaluate_clear(circ, q, noise=nm, rng=np.random.default_rng(2))
    assert np.corrcoef(noisy.ravel(), clear.ravel())[0, 1] > 0.9
