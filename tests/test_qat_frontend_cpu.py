"""QAT front-end and the unmodified-reference-script harness, CPU side.

The Brevitas layers come from the import shim in dct-cryptonets_b200/compat/brevitas (Brevitas is not installable here); the
front-end must take every quantiser (input, weights, activations) from the modules instead of calibrating it."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "dct-cryptonets_b200", "compat")
if COMPAT not in sys.path:
    sys.path.append(COMPAT)

import brevitas.nn as qnn                                   # noqa: E402  (the shim, unless a real Brevitas is installed)
from brevitas.quant import Int8ActPerTensorFloat, Int8WeightPerTensorFloat   # noqa: E402
from tfx_b200 import circuit as C                           # noqa: E402

REF = "/root/reference/dct-cryptonets"
QCONV = dict(weight_bit_width=4, weight_quant=Int8WeightPerTensorFloat, bias=False, bias_quant=None, narrow_range=True)
QID = dict(bit_width=4, act_quant=Int8ActPerTensorFloat)


def _init(net):
    torch.manual_seed(0)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    return net.double().eval()


def test_shim_quantisers():
    conv = qnn.QuantConv2d(3, 5, 3, **QCONV)
    w, s = conv.tfx_weight_quant()
    assert int(w.abs().max()) == 7 and w.min() >= -7                       # narrow range: [-7, 7]
    assert abs(s - float(conv.weight.detach().abs().max()) / 7) < 1e-7
    assert torch.allclose(conv.quant_weight(), w.float() * s)
    ident = qnn.QuantIdentity(return_quant_tensor=False, **QID)
    assert ident.tfx_act_quant() == (1.0 / 8, -8, 7)                        # untrained threshold 1.0, signed 4 bits
    relu = qnn.QuantReLU(bit_width=4)
    assert relu.tfx_act_quant() == (pytest.approx(1.0 / 15, rel=1e-6), 0, 15)
    x = torch.linspace(-2, 2, 41) + 0.013                                   # keeps x / scale away from rounding ties
    assert torch.equal(ident(x).detach(), torch.clamp(torch.round(x * 8), -8, 7) / 8)
    assert torch.allclose(relu(x).detach(), torch.clamp(torch.round(torch.relu(x) * 15), 0, 15) / 15, atol=1e-6)
    keys = list(ident.state_dict().keys())
    assert keys == ["act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value"]
    with pytest.raises(NotImplementedError):
        qnn.QuantIdentity(return_quant_tensor=True, **QID)


def test_qat_chain_equals_fake_quant_model():
    """no residual add, all accumulator bits kept: the integer circuit IS the fake-quant float model"""
    net = _init(nn.Sequential(qnn.QuantIdentity(return_quant_tensor=False, **QID), qnn.QuantConv2d(3, 6, 3, padding=1, **QCONV), nn.BatchNorm2d(6),
                              qnn.QuantReLU(bit_width=4), qnn.QuantConv2d(6, 4, 3, padding=1, **QCONV), nn.BatchNorm2d(4),
                              qnn.QuantIdentity(return_quant_tensor=False, **QID)))
    calib = torch.randn(64, 3, 6, 6, dtype=torch.float64) * 0.6
    circ = C.build_circuit(net, calib, n_bits=5, rounding_threshold_bits=16, p_error=0.01)
    assert (circ.input_quant.scale, circ.input_quant.qmin, circ.input_quant.qmax) == (0.125, -8, 7)      # from the module, not n_bits
    convs = [op for op in circ.ops if op.kind == "conv"]
    assert all(np.abs(op.raw_weight).max() == 7 for op in convs)
    tl = circ.lookups()
    assert (tl[0].out.scale, tl[0].out.qmin, tl[0].out.qmax) == (1.0 / 15, 0, 15)
    assert (tl[1].out.scale, tl[1].out.qmin, tl[1].out.qmax) == (0.125, -8, 7)
    x = calib[:16]
    with torch.no_grad():
        ref = net(x).numpy()
    out = C.dequantize_output(circ, C.evaluate_clear(circ, C.quantize_input(circ, x.numpy()))).reshape(ref.shape)
    assert np.array_equal(np.rint(ref * 8), np.rint(out * 8))


def test_qat_residual_block_tracks_fake_quant_model():
    """residual add: one operand is re-expressed on the other's grid (what Concrete-ML's QuantizedAdd does with a table per
    operand); outputs stay within one output level of the float fake-quant model"""
    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.inp = qnn.QuantIdentity(return_quant_tensor=False, **QID)
            self.c0 = qnn.QuantConv2d(3, 6, 1, **QCONV); self.b0 = nn.BatchNorm2d(6); self.r0 = qnn.QuantReLU(bit_width=4)
            self.C1 = qnn.QuantConv2d(6, 6, 3, padding=1, **QCONV); self.BN1 = nn.BatchNorm2d(6); self.relu1 = qnn.QuantReLU(bit_width=4)
            self.C2 = qnn.QuantConv2d(6, 6, 3, padding=1, **QCONV); self.BN2 = nn.BatchNorm2d(6)
            self.quant_out = qnn.QuantIdentity(return_quant_tensor=False, scaling_init=1.0, **QID)
            self.relu2 = qnn.QuantReLU(bit_width=4)

        def forward(self, x):
            x = self.r0(self.b0(self.c0(self.inp(x))))
            out = self.quant_out(self.BN2(self.C2(self.relu1(self.BN1(self.C1(x))))))
            return self.relu2(torch.add(out, x))

    net = _init(Block())
    calib = torch.randn(64, 3, 6, 6, dtype=torch.float64) * 0.6
    circ = C.build_circuit(net, calib, n_bits=5, rounding_threshold_bits=16, p_error=0.01)
    x = calib[:16]
    with torch.no_grad():
        ref = net(x).numpy()
    out = C.dequantize_output(circ, C.evaluate_clear(circ, C.quantize_input(circ, x.numpy()))).reshape(ref.shape)
    lvl = np.abs(np.rint(ref * 15) - np.rint(out * 15))
    assert lvl.max() <= 1 and (lvl > 0).mean() < 0.25


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_unmodified_reference_script_runs_to_keygen(tmp_path):
    """tools/run_reference_eval.py drives the reference's untouched homomorphic_eval.py: its own data pipeline and QAT model,
    this backend's compile_brevitas_qat_model / Configuration / graph.maximum_integer_bit_width() / .mlir; without a GPU it must
    stop at fhe_circuit.keygen() (no CPU fallback), with a GPU it must finish."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_eval.py"), "--workdir", str(tmp_path), "--synthetic-cifar", "120", "--",
           "--dataset", "cifar10", "--model", "ResNet20qat", "--dct_status", "--channels", "24", "--filter_size", "4", "--image_size_dct", "16",
           "--bit_width", "4", "--fhe_mode", "simulate", "--calib_batch_size", "100", "--test_batch_size", "2", "--test_subset", "2",
           "--rounding_threshold_bits", "6", "--n_bits", "5", "--p_error", "0.01"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    log = r.stdout + r.stderr
    assert "Time for FHE compilation" in log and "it works in FHE!!" in log, log[-2000:]
    assert os.path.getsize(tmp_path / "mlir.txt") > 1000
    if torch.cuda.is_available():
        assert r.returncode == 0 and "Done" in log, log[-2000:]
    else:
        assert r.returncode != 0 and "keygen" in log and "no CPU fallback" in log, log[-2000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_checkpoint_format_loads_and_changes_the_circuit(tmp_path):
    """A checkpoint in the reference's own container format (train.py:83-89: {'epoch', 'state' = DataParallel state dict, 'prec1',
    'prec5', 'optimizer'}) is picked up by the unmodified driver (homomorphic_eval.py:247-253) and its weights — not the random
    initialisation — are what gets compiled: the circuit text written to mlir.txt differs from the no-checkpoint run.
    (Brevitas itself is not installable here: parameter names are those of compat/brevitas, see DESIGN.md.)"""
    pkg = os.path.join(ROOT, "dct-cryptonets_b200")
    ckpt = tmp_path / "best.tar"
    make = f"""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, {REF!r}); sys.path.insert(0, {pkg!r}); sys.path.append({os.path.join(pkg, "compat")!r})
os.environ["BREVITAS_IGNORE_MISSING_KEYS"] = "1"
from io_utils import model_dict
from utils import BaselineTrain
torch.manual_seed(7)
model = nn.DataParallel(BaselineTrain(model_dict["ResNet20qat"](bit_width=4, in_channels=24, img_size=16), 10))
with torch.no_grad():
    for p_ in model.parameters():
        p_.add_(0.05 * torch.randn_like(p_))          # "trained": anything but the seeded initialisation
torch.save({{"epoch": 3, "state": model.state_dict(), "prec1": 91.25, "prec5": 99.5, "optimizer": {{}}}}, {str(ckpt)!r})
"""
    r = subprocess.run([sys.executable, "-c", make], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and ckpt.exists(), r.stderr[-2000:]
    texts = {}
    for tag, extra in (("ckpt", ["--checkpoint_path", str(ckpt)]), ("none", [])):
        wd = tmp_path / tag
        cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_eval.py"), "--workdir", str(wd), "--synthetic-cifar", "120", "--",
               "--dataset", "cifar10", "--model", "ResNet20qat", "--dct_status", "--channels", "24", "--filter_size", "4", "--image_size_dct", "16",
               "--bit_width", "4", "--fhe_mode", "simulate", "--calib_batch_size", "100", "--test_batch_size", "2", "--test_subset", "2",
               "--rounding_threshold_bits", "6", "--n_bits", "5", "--p_error", "0.01"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        log = r.stdout + r.stderr
        assert "Time for FHE compilation" in log, log[-2000:]
        if tag == "ckpt":
            assert "Loaded checkpoint" in log and "91.250% Top-1 Acc. @ epoch 3" in log and "No checkpoint loaded" not in log, log[-2000:]
        else:
            assert "No checkpoint loaded" in log
        texts[tag] = open(wd / "mlir.txt").read()
    assert texts["ckpt"] != texts["none"]

