"""Opt-in circuit variants (not the default layout): they must stay word-for-word equal to the oracle like everything else.
(Round 1 shipped this test as xfail(strict=False) before its first GPU run; it passed on the B200 and is strict now.)"""
import numpy as np
import pytest
import torch
import torch.nn as nn

from tfx_b200 import circuit as C
from tfx_b200.executor import CircuitExecutor
from tfx_b200.resnet_dct import ResidualBlock
from test_circuit_gpu import TOY_BIT, TOY_TLU

pytestmark = pytest.mark.gpu


def test_fused_residual_lookups_match_oracle_and_clear(gpu_ctx, oracle):
    """opt-in fused residual lookups (circuit.FusedAddOp): GPU ciphertexts == oracle circuit word for word, decrypted == clear"""
    from oracle import circuit_oracle as CO
    torch.manual_seed(4)
    net = nn.Sequential(nn.Conv2d(3, 5, 1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), ResidualBlock(5, 5, False),
                        ResidualBlock(5, 7, True), nn.AvgPool2d(2), nn.Flatten()).eval()
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data = torch.linspace(0.1, 0.2, m.num_features)
            m.bias.data = torch.linspace(-0.5, 0.5, m.num_features)
    calib = torch.randn(32, 3, 4, 4)
    circ = C.build_circuit(net, calib, n_bits=5, rounding_threshold_bits=6, p_error=0.01, fuse_residual=True)
    assert any(op.kind == "fadd" for op in circ.ops)
    ex = CircuitExecutor(circ, (TOY_TLU, TOY_BIT), ctx=gpu_ctx, input_std=2.0**-50)
    ex.keygen(seed=15)
    q_in = C.quantize_input(circ, calib[:1].numpy())[0]
    got = gpu_ctx.to_host_u64(ex.run(ex.encrypt(q_in, enc_seed=16)))
    keys = CO.OracleKeys((TOY_TLU, TOY_BIT), 15)
    ref = CO.run_circuit(circ, keys, CO.encrypt_input(circ, keys, q_in, 2.0**-50, 16))
    assert np.array_equal(got, ref)
    clear = C.evaluate_clear(circ, q_in[None])[0]
    assert np.array_equal(CO.decrypt_output(circ, keys, ref).reshape(clear.shape), clear)
