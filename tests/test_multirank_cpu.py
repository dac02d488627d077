"""world_size-2 gloo test of the N > 1 host logic: per-layer channel partition + all-gather reassembly, driven with
the clear integer semantics of the circuit (the CUDA kernels are not involved; they are covered by -m gpu tests)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tfx_b200.executor import channel_range, gather_channels


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn as nn
        from tfx_b200 import circuit as C
        from tfx_b200.resnet_dct import ResidualBlock
        torch.manual_seed(0)
        net = nn.Sequential(nn.Conv2d(3, 5, 1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), ResidualBlock(5, 5, False),
                            ResidualBlock(5, 7, True), nn.AvgPool2d(2), nn.Flatten()).eval()
        calib = torch.randn(32, 3, 4, 4)
        circ = C.build_circuit(net, calib, 5, 6, 0.01)
        q = C.quantize_input(circ, calib[:1].numpy())
        want = C.evaluate_clear(circ, q)
        # sharded evaluation: same walk as CircuitExecutor.run, integers standing in for ciphertext rows
        vals = {circ.input_id: q[0]}
        acc_local = {}
        for op in circ.ops:
            if op.kind == "conv":
                Cc = op.out_shape[0]
                lo, hi, per = (0, Cc, Cc) if op.dst == circ.output_id else channel_range(Cc, rank, world)
                full = C._int_conv(vals[op.src][None], op, op.raw_weight)[0]
                acc_local[op.dst] = (full[lo:hi], lo, hi, per)
                if op.dst == circ.output_id:
                    vals[op.dst] = full
            elif op.kind == "add":
                Cc = op.shape[0]
                lo, hi, per = channel_range(Cc, rank, world)
                acc_local[op.dst] = (vals[op.a][lo:hi] + vals[op.b][lo:hi], lo, hi, per)
            else:
                acc, lo, hi, per = acc_local.pop(op.src)
                Cc, H, W = op.shape
                lin = next(o for o in circ.ops if getattr(o, "dst", None) == op.src)
                sub = C.TluOp(op.name, op.src, op.dst, (hi - lo, H, W), op.acc_bits, op.keep_bits, op.tables[lo:hi], op.out,
                              chan_bits=None if op.chan_bits is None else op.chan_bits[lo:hi])
                out = C.tlu_apply(sub, C.channel_offsets(lin.offset, Cc)[lo:hi], acc[None])[0] if hi > lo else np.zeros((0, H, W), np.int64)
                local = torch.from_numpy(out.reshape(-1, 1).astype(np.int64))
                full = gather_channels(local, Cc, per, H * W, world)
                vals[op.dst] = full.numpy().reshape(Cc, H, W)
        got = vals[circ.output_id].reshape(want.shape[1:])
        ret[rank] = bool(np.array_equal(got, want[0]))
    finally:
        dist.destroy_process_group()


def test_channel_range_covers_everything():
    for C_ in (48, 56, 64, 7, 3):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = channel_range(C_, r, world)
                assert 0 <= lo <= hi <= C_ and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(C_))


def test_two_rank_partition_and_gather_reproduce_single_rank():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
