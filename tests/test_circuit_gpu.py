"""End-to-end GPU parity: a compiled circuit run by the CUDA executor against the CPU oracle (every output ciphertext
word) and against the clear integer evaluator (decrypted values, tight-noise toy parameters)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from tfx_b200 import circuit as C
from tfx_b200.binding import PbsParams
from tfx_b200.executor import CircuitExecutor, RunStats
from tfx_b200.resnet_dct import ResidualBlock

pytestmark = pytest.mark.gpu

# insecure toy sets with negligible noise: outputs must equal the clear evaluation exactly
TOY_TLU = PbsParams(n=96, k=1, N=2048, bsk_base_log=12, bsk_level=3, ksk_base_log=4, ksk_level=6, lwe_std=2.0**-40, glwe_std=2.0**-55)
TOY_BIT = PbsParams(n=80, k=2, N=512, bsk_base_log=12, bsk_level=3, ksk_base_log=4, ksk_level=6, lwe_std=2.0**-40, glwe_std=2.0**-55)


class TinyNet(nn.Module):
    def __init__(self, cin=3, c1=4, c2=6, avg=2):
        super().__init__()
        self.trunk = nn.Sequential(nn.Conv2d(cin, c1, 1, bias=False), nn.BatchNorm2d(c1), nn.ReLU(),
                                   ResidualBlock(c1, c1, False), ResidualBlock(c1, c2, True),
                                   nn.AvgPool2d(avg), nn.Flatten())
        self.final_feat_dim = c2

    def forward(self, x):
        return self.trunk(x)


def build(seed=0):
    torch.manual_seed(seed)
    model = TinyNet().eval()
    calib = torch.randn(40, 3, 4, 4)
    circ = C.build_circuit(model, calib, n_bits=5, rounding_threshold_bits=6, p_error=0.01)
    return model, calib, circ


def test_tiny_circuit_matches_oracle_and_clear(gpu_ctx, oracle):
    from oracle import circuit_oracle as CO
    model, calib, circ = build()
    assert any(op.lsbs > 0 for op in circ.lookups()) and any(op.kind == "add" for op in circ.ops)
    ex = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    ex.keygen(seed=5)
    q_in = C.quantize_input(circ, calib[:1].numpy())[0]
    cts = ex.encrypt(q_in, enc_seed=6)
    stats = RunStats()
    out = ex.run(cts, stats)
    got = gpu_ctx.to_host_u64(out)
    # CPU oracle with its own keygen from the same seed
    keys = CO.OracleKeys((TOY_TLU, TOY_BIT), 5)
    o_cts = CO.encrypt_input(circ, keys, q_in, 2.0**-50, 6)
    assert np.array_equal(gpu_ctx.to_host_u64(cts), o_cts)
    ref = CO.run_circuit(circ, keys, o_cts)
    assert np.array_equal(got, ref), "GPU circuit output ciphertexts differ from the oracle's"
    clear = C.evaluate_clear(circ, q_in[None])[0]
    assert np.array_equal(ex.decrypt(out).reshape(clear.shape), clear)
    assert np.array_equal(CO.decrypt_output(circ, keys, ref).reshape(clear.shape), clear)
    cnt = circ.pbs_count()
    assert stats.pbs_tlu == cnt["tlu"] and stats.pbs_bit == cnt["bit"] and stats.launches > 0


def test_per_channel_widths_match_oracle_and_clear(gpu_ctx, oracle):
    """opt-in per-channel accumulator widths (width-sorted rows, one table-lookup keyswitch per width, gather / scatter): GPU
    ciphertexts == oracle circuit word for word, decrypted == clear evaluator; also with the two-stream split forced on"""
    from oracle import circuit_oracle as CO
    torch.manual_seed(4)
    net = nn.Sequential(nn.Conv2d(3, 5, 1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), ResidualBlock(5, 5, False),
                        ResidualBlock(5, 7, True), nn.AvgPool2d(2), nn.Flatten()).eval()
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data = torch.linspace(0.2, 2.0, m.num_features)
            m.bias.data = torch.linspace(-0.5, 0.5, m.num_features)
    calib = torch.randn(32, 3, 4, 4)
    circ = C.build_circuit(net, calib, n_bits=5, rounding_threshold_bits=6, p_error=0.01, per_channel_widths=True)
    assert any(op.chan_bits is not None and len(set(op.chan_bits.tolist())) > 1 for op in circ.lookups())
    ex = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    ex.keygen(seed=13)
    q_in = C.quantize_input(circ, calib[:1].numpy())[0]
    cts = ex.encrypt(q_in, enc_seed=14)
    stats = RunStats()
    got = gpu_ctx.to_host_u64(ex.run(cts, stats))
    keys = CO.OracleKeys((TOY_TLU, TOY_BIT), 13)
    ref = CO.run_circuit(circ, keys, CO.encrypt_input(circ, keys, q_in, 2.0**-50, 14))
    assert np.array_equal(got, ref)
    clear = C.evaluate_clear(circ, q_in[None])[0]
    assert np.array_equal(CO.decrypt_output(circ, keys, ref).reshape(clear.shape), clear)
    assert stats.pbs_bit == circ.pbs_count()["bit"]
    ex.max_chains, ex.WAVE_ROWS = 3, 8                       # concurrent wave-sized chains forced on (toy wave size)
    assert np.array_equal(gpu_ctx.to_host_u64(ex.run(cts)), ref)


def test_two_stream_lookup_layers_give_identical_ciphertexts(gpu_ctx):
    """the multi-GPU executor cuts a rank's share of every lookup layer into wave-sized chunks whose chains run on several
    streams (executor.py); forced on here on one GPU with a toy wave size: every output word must equal the single-stream
    run, repeatedly (a race would show up as a difference)"""
    model, calib, circ = build(seed=2)
    one = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    one.max_chains = 1
    one.keygen(seed=7)
    two = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    two.max_chains, two.WAVE_ROWS = 4, 8
    two.use_keys(one.keys)
    for img in range(3):
        q_in = C.quantize_input(circ, calib[img:img + 1].numpy())[0]
        cts = one.encrypt(q_in, enc_seed=8 + img)
        want = gpu_ctx.to_host_u64(one.run(cts))
        for _ in range(2):
            assert np.array_equal(gpu_ctx.to_host_u64(two.run(cts)), want)
    assert len(two._sides) >= 2 and not one._sides


def test_qat_circuit_executes_like_the_clear_evaluator(gpu_ctx):
    """a Brevitas-style QAT block (quantisers taken from the modules) through the CUDA executor: decrypted outputs
    equal the clear integer evaluator under tight-noise parameters"""
    import os, sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dct-cryptonets_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)
    import brevitas.nn as qnn
    from brevitas.quant import Int8ActPerTensorFloat, Int8WeightPerTensorFloat
    qconv = dict(weight_bit_width=4, weight_quant=Int8WeightPerTensorFloat, bias=False, bias_quant=None, narrow_range=True)
    qid = dict(bit_width=4, act_quant=Int8ActPerTensorFloat)

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.inp = qnn.QuantIdentity(return_quant_tensor=False, **qid)
            self.c0 = qnn.QuantConv2d(3, 4, 1, **qconv); self.b0 = nn.BatchNorm2d(4); self.r0 = qnn.QuantReLU(bit_width=4)
            self.C1 = qnn.QuantConv2d(4, 4, 3, padding=1, **qconv); self.BN1 = nn.BatchNorm2d(4); self.relu1 = qnn.QuantReLU(bit_width=4)
            self.C2 = qnn.QuantConv2d(4, 4, 3, padding=1, **qconv); self.BN2 = nn.BatchNorm2d(4)
            self.quant_out = qnn.QuantIdentity(return_quant_tensor=False, scaling_init=1.0, **qid)
            self.relu2 = qnn.QuantReLU(bit_width=4)
            self.pool = nn.AvgPool2d(2); self.q = qnn.QuantIdentity(return_quant_tensor=False, **qid); self.flat = nn.Flatten()

        def forward(self, x):
            x = self.r0(self.b0(self.c0(self.inp(x))))
            out = self.quant_out(self.BN2(self.C2(self.relu1(self.BN1(self.C1(x))))))
            return self.flat(self.q(self.pool(self.relu2(torch.add(out, x)))))

    torch.manual_seed(3)
    net = Block().eval()
    calib = torch.randn(40, 3, 4, 4) * 0.6
    circ = C.build_circuit(net, calib, n_bits=5, rounding_threshold_bits=6, p_error=0.01)
    assert circ.input_quant.scale == 0.125 and not circ.output_is_acc
    ex = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    ex.keygen(seed=11)
    for img in range(2):
        q_in = C.quantize_input(circ, calib[img:img + 1].numpy())[0]
        out = ex.run(ex.encrypt(q_in, enc_seed=12 + img))
        clear = C.evaluate_clear(circ, q_in[None])[0]
        assert np.array_equal(ex.decrypt(out).reshape(clear.shape), clear)


def test_quantized_module_execute_matches_simulate(gpu_ctx):
    from tfx_b200.quantized_module import QuantizedModule
    torch.manual_seed(1)
    model = TinyNet().eval()
    calib = torch.randn(40, 3, 4, 4)
    qm = QuantizedModule.compile(model, calib, 5, 6, 0.01, params=(TOY_TLU, TOY_BIT))
    qm.fhe_circuit.keygen(seed=9)
    x = calib[:2].numpy()
    y_exec = qm.forward(x, fhe="execute")
    y_sim = qm.forward(x, fhe="simulate")
    assert y_exec.shape == (2, 6) and np.array_equal(y_exec, y_sim)


def test_client_server_deployment_roundtrip(gpu_ctx, tmp_path):
    """FHEModelDev.save -> FHEModelClient (keygen, encrypt) -> bytes -> FHEModelServer.run -> bytes -> client decrypt"""
    from concrete.ml.deployment import FHEModelClient, FHEModelDev, FHEModelServer
    from tfx_b200.quantized_module import QuantizedModule
    torch.manual_seed(2)
    model = TinyNet().eval()
    calib = torch.randn(40, 3, 4, 4)
    qm = QuantizedModule.compile(model, calib, 5, 6, 0.01, params=(TOY_TLU, TOY_BIT))
    FHEModelDev(str(tmp_path), qm).save()
    client = FHEModelClient(str(tmp_path))
    client.generate_private_and_evaluation_keys()
    eval_keys = client.get_serialized_evaluation_keys()
    x = calib[:1].numpy()
    blob = client.quantize_encrypt_serialize(x)
    server = FHEModelServer(str(tmp_path))
    result = server.run(blob, eval_keys)
    assert isinstance(result, bytes)
    y = client.deserialize_decrypt_dequantize(result)
    assert np.array_equal(y, qm.forward(x, fhe="simulate"))
    assert server._ex.keys is not client._ex.keys


def test_approximate_rounding_matches_oracle_and_clear(gpu_ctx, oracle):
    """rounding method 'approximate' (reference README.md:96-113): no bit-extraction chain, same table; with negligible
    noise the thresholds are exact, so outputs equal the clear evaluator; ciphertexts equal the oracle's."""
    from oracle import circuit_oracle as CO
    torch.manual_seed(0)
    model = TinyNet().eval()
    calib = torch.randn(40, 3, 4, 4)
    circ = C.build_circuit(model, calib, 5, 6, 0.01, rounding_method="approximate")
    assert circ.pbs_count()["bit"] == 0
    ex = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    ex.keygen(seed=5)
    q_in = C.quantize_input(circ, calib[:1].numpy())[0]
    stats = RunStats()
    out = ex.run(ex.encrypt(q_in, enc_seed=6), stats)
    keys = CO.OracleKeys((TOY_TLU, TOY_BIT), 5)
    ref = CO.run_circuit(circ, keys, CO.encrypt_input(circ, keys, q_in, 2.0**-50, 6))
    assert np.array_equal(gpu_ctx.to_host_u64(out), ref), "GPU ciphertexts differ from the oracle's (approximate mode)"
    assert stats.pbs_bit == 0 and stats.pbs_tlu == circ.pbs_count()["tlu"]
    clear = C.evaluate_clear(circ, q_in[None])[0].astype(np.float64)
    dec = ex.decrypt(out).reshape(clear.shape).astype(np.float64)
    # approximate rounding: the mod-switch noise blurs every rounding threshold (that is the approximation), so the
    # decrypted features track the clear evaluation without being equal to it
    assert np.abs(dec - clear).max() <= 0.35 * max(1.0, np.abs(clear).max()), (dec, clear)


def test_maxpool_stem_matches_oracle_and_clear(gpu_ctx, oracle):
    """MaxPool2d stem (reference models/backbone.py:153-160) as chained relu lookups over window taps (circuit.LinOp): GPU ciphertexts
    == oracle circuit word for word, decrypted == clear evaluator == torch's max_pool2d on the quantised activations"""
    from oracle import circuit_oracle as CO
    torch.manual_seed(6)
    net = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1, bias=False), nn.BatchNorm2d(4), nn.ReLU(), nn.MaxPool2d(3, stride=2, padding=1),
                        nn.Conv2d(4, 5, 1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), nn.AvgPool2d(2), nn.Flatten()).eval()
    calib = torch.randn(32, 3, 6, 6)
    circ = C.build_circuit(net, calib, n_bits=5, rounding_threshold_bits=6, p_error=0.01)
    assert sum(op.kind == "lin" for op in circ.ops) == 9
    ex = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    ex.keygen(seed=17)
    q_in = C.quantize_input(circ, calib[:1].numpy())[0]
    got = gpu_ctx.to_host_u64(ex.run(ex.encrypt(q_in, enc_seed=18)))
    keys = CO.OracleKeys((TOY_TLU, TOY_BIT), 17)
    ref = CO.run_circuit(circ, keys, CO.encrypt_input(circ, keys, q_in, 2.0**-50, 18))
    assert np.array_equal(got, ref)
    clear = C.evaluate_clear(circ, q_in[None])[0]
    assert np.array_equal(ex.decrypt(ex.run(ex.encrypt(q_in, enc_seed=18))).reshape(clear.shape), clear)

