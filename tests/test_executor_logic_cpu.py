"""The CUDA executor's HOST logic (layer loop, rounding chains, per-channel offsets and widths, width-sorted rows, rank
partition) exercised on the CPU: `CircuitExecutor.run` is driven with a stand-in context / key set whose primitives are the
oracle's (keyswitch, PBS, conv, axpby on CPU tensors), and its output ciphertexts must equal oracle/circuit_oracle.run_circuit
word for word.  This checks everything in executor.py except the kernels themselves (their parity is tests/test_kernels_gpu.py)
— test infrastructure only; the product path never sees these stand-ins."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from tfx_b200 import circuit as C
from tfx_b200.binding import PbsParams
from tfx_b200.executor import CircuitExecutor, RunStats
from tfx_b200.resnet_dct import ResidualBlock

TLU = PbsParams(n=64, k=1, N=2048, bsk_base_log=12, bsk_level=3, ksk_base_log=4, ksk_level=6, lwe_std=2.0**-40, glwe_std=2.0**-55)
BIT = PbsParams(n=48, k=2, N=512, bsk_base_log=12, bsk_level=3, ksk_base_log=4, ksk_level=6, lwe_std=2.0**-40, glwe_std=2.0**-55)


def _np(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().numpy().view(np.uint64)


def _pt(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64))


class FakeContext:
    """Context look-alike on CPU tensors (int64 views of torus words), leveled ops by the oracle"""
    device = torch.device("cpu")

    def __init__(self, O):
        self.O = O

    def empty_u64(self, *shape):
        return torch.empty(*shape, dtype=torch.int64)

    def to_device_u64(self, a):
        return _pt(np.ascontiguousarray(a, dtype=np.uint64))

    def to_host_u64(self, t):
        return _np(t)

    def conv2d(self, x, w, stride, pad, bias_pt=None, oc_range=None, depthwise=False, out=None):
        full = self.O.conv2d(_np(x), w.numpy(), stride, pad, None if bias_pt is None else _np(bias_pt), depthwise=depthwise)
        ob, oe = (0, full.shape[0]) if oc_range is None else oc_range
        return _pt(full[ob:oe])

    def axpby(self, a, sa, b=None, sb=0, body_const=0, out=None):
        r = _pt(self.O.axpby(_np(a), sa, None if b is None else _np(b), sb, body_const))
        if out is not None:
            out.copy_(r.view(out.shape))
            return out
        return r.view(a.shape)


class FakeKeys:
    """KeySet look-alike: keyswitch / PBS by the oracle with the oracle's own keys"""

    def __init__(self, O, okeys, params):
        self.O, self.k, self.params = O, okeys, list(params)
        self.big_dim = okeys.big_dim

    def keyswitch(self, set_id, cts, shift=0, body_offset=0, out=None, ctx=None):
        p = self.params[set_id]
        r = self.O.keyswitch(self.k.ksk[set_id], _np(cts), p.ksk_base_log, p.ksk_level, shift=shift, body_offset=body_offset)
        if out is not None:
            out.copy_(_pt(r))
            return out
        return _pt(r)

    def pbs(self, set_id, cts, luts, lut_index, mode=0, body_const=0, out=None, ctx=None):
        p = self.params[set_id]
        prev = None if out is None else _np(out).copy()
        r = self.O.pbs(self.k.bsk_f[set_id], p.bsk_base_log, _np(cts), _np(luts), lut_index.numpy().astype(np.uint32), mode=mode,
                       body_const=body_const, out=prev, big_dim=self.big_dim)
        if out is not None:
            out.copy_(_pt(r))
            return out
        return _pt(r)


def _net(spread: bool = True):
    torch.manual_seed(4)
    net = nn.Sequential(nn.Conv2d(3, 5, 1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), ResidualBlock(5, 5, False),
                        ResidualBlock(5, 7, True), nn.AvgPool2d(2), nn.Flatten()).eval()
    for m in net.modules():                                  # uneven channel ranges so that per-channel widths really differ
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data = torch.linspace(0.2, 2.0, m.num_features) if spread else torch.linspace(0.1, 0.2, m.num_features)
            m.bias.data = torch.linspace(-0.5, 0.5, m.num_features)
    return net, torch.randn(32, 3, 4, 4)


def _maxpool_net():
    """conv stem + ReLU + MaxPool2d(3, 2, 1) (the reference's RGB-224 stem shape, models/backbone.py:153-160) + 1x1 conv + pool"""
    torch.manual_seed(6)
    net = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1, bias=False), nn.BatchNorm2d(4), nn.ReLU(), nn.MaxPool2d(3, stride=2, padding=1),
                        nn.Conv2d(4, 5, 1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), nn.AvgPool2d(2), nn.Flatten()).eval()
    return net, torch.randn(32, 3, 6, 6)


@pytest.mark.parametrize("mode", ["tensor_wide", "per_channel_offsets", "per_channel_widths", "approximate_widths", "fused_widths",
                                  "maxpool_widths"])
def test_executor_host_logic_equals_oracle_circuit(oracle, mode):
    from oracle import circuit_oracle as CO
    net, calib = _maxpool_net() if mode.startswith("maxpool") else _net(spread=not mode.startswith("fused"))   # small BatchNorm gains: the shortcut's integer weight m_c is >= 8
    circ = C.build_circuit(net, calib, n_bits=5, rounding_threshold_bits=6, p_error=0.01,
                           per_channel_offsets=(mode != "tensor_wide"), per_channel_widths=mode.endswith("widths"),
                           rounding_method="approximate" if mode.startswith("approximate") else "exact",
                           fuse_residual=mode.startswith("fused"))
    assert any(op.kind == "fadd" for op in circ.ops) == mode.startswith("fused")
    assert any(op.kind == "lin" for op in circ.ops) == mode.startswith("maxpool")
    if mode.startswith("maxpool"):
        return _check_maxpool_circuit(oracle, CO, circ, net, calib)
    if mode.endswith("widths"):
        assert any(op.chan_bits is not None and len(set(op.chan_bits.tolist())) > 1 for op in circ.lookups())
    okeys = CO.OracleKeys((TLU, BIT), 9)
    q = C.quantize_input(circ, calib[:1].numpy())[0]
    o_cts = CO.encrypt_input(circ, okeys, q, 2.0**-50, 10)
    want = CO.run_circuit(circ, okeys, o_cts)
    if circ.rounding_method == "exact":                      # (approximate rounding is allowed to differ from the clear model by design)
        clear = C.evaluate_clear(circ, q[None])[0].reshape(-1)
        assert np.array_equal(CO.decrypt_output(circ, okeys, want), clear)
    for world in (1, 2):
        outs = []
        for rank in range(world):
            ex = CircuitExecutor(circ, (TLU, BIT), ctx=FakeContext(oracle), rank=rank, world_size=world, input_std=2.0**-50)
            ex.max_chains = 1
            ex.use_keys(FakeKeys(oracle, okeys, (TLU, BIT)))
            if world > 1:                                    # stand-in for the all-gather: every rank's block, computed here in turn
                ex._gather = None
            outs.append(ex)
        if world == 1:
            stats = RunStats()
            got = _np(outs[0].run(_pt(o_cts), stats))
            assert np.array_equal(got, want)
            cnt = circ.pbs_count()
            assert stats.pbs_tlu == cnt["tlu"] and stats.pbs_bit == cnt["bit"]
        else:
            got = _run_two_ranks(outs, o_cts)
            assert np.array_equal(got, want)


def _check_maxpool_circuit(oracle, CO, circ, net, calib):
    """MaxPool2d as chained b + relu(a - b) lookups over window taps: exact in the clear, executor == oracle evaluator word for word
    (1 and 2 ranks), decrypted == clear"""
    q = C.quantize_input(circ, calib[:4].numpy())
    col = {}
    C.evaluate_clear(circ, q, collect=col)
    src = circ.ops[1]                                          # the stem's ReLU lookup
    top = [op for op in circ.ops if op.kind == "lin"][-1]
    ref = torch.nn.functional.max_pool2d(torch.from_numpy(col[src.dst].astype(np.float64)), 3, 2, 1).numpy().astype(np.int64)
    assert np.array_equal(col[top.dst], ref)
    c2 = C.circuit_from_portable(*C.circuit_to_portable(circ))
    assert c2.to_text() == circ.to_text()
    okeys = CO.OracleKeys((TLU, BIT), 9)
    o_cts = CO.encrypt_input(circ, okeys, q[0], 2.0**-50, 10)
    want = CO.run_circuit(circ, okeys, o_cts)
    assert np.array_equal(CO.decrypt_output(circ, okeys, want), C.evaluate_clear(circ, q[:1])[0].reshape(-1))
    execs = []
    for world in (1, 2):
        outs = []
        for rank in range(world):
            ex = CircuitExecutor(circ, (TLU, BIT), ctx=FakeContext(oracle), rank=rank, world_size=world, input_std=2.0**-50)
            ex.max_chains = 1
            ex.use_keys(FakeKeys(oracle, okeys, (TLU, BIT)))
            outs.append(ex)
        got = _np(outs[0].run(_pt(o_cts))) if world == 1 else _run_two_ranks(outs, o_cts)
        assert np.array_equal(got, want)


def _run_two_ranks(execs, o_cts):
    """runs the ranks in lock step on one process: each lookup layer's local blocks are concatenated in place of the all-gather"""
    import threading
    world = len(execs)
    barrier = threading.Barrier(world)
    blocks, results = {}, [None] * world
    lock = threading.Lock()

    def gather_for(rank):
        def gather(local, Cc, per, hw):
            with lock:
                blocks[rank] = local
            barrier.wait()
            full = torch.cat([blocks[r] for r in range(world)], dim=0)[: Cc * hw]
            barrier.wait()
            return full
        return gather

    def work(rank):
        execs[rank]._gather = gather_for(rank)
        results[rank] = _np(execs[rank].run(_pt(o_cts)))

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert np.array_equal(results[0], results[1])
    return results[0]
