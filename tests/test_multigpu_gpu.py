"""N-GPU run == 1-GPU run, word for word.  Needs >= 2 visible GPUs; skipped otherwise."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import torch.nn as nn
        from tfx_b200 import circuit as C
        from tfx_b200.binding import Context, PbsParams
        from tfx_b200.executor import CircuitExecutor
        from tfx_b200.resnet_dct import ResidualBlock
        torch.manual_seed(0)
        net = nn.Sequential(nn.Conv2d(3, 5, 1, bias=False), nn.BatchNorm2d(5), nn.ReLU(), ResidualBlock(5, 5, False),
                            ResidualBlock(5, 7, True), nn.AvgPool2d(2), nn.Flatten()).eval()
        calib = torch.randn(32, 3, 4, 4)
        circ = C.build_circuit(net, calib, 5, 6, 0.01)
        tlu = PbsParams(n=96, k=1, N=2048, bsk_base_log=12, bsk_level=3, ksk_base_log=4, ksk_level=6, lwe_std=2.0**-40, glwe_std=2.0**-55)
        bit = PbsParams(n=80, k=2, N=512, bsk_base_log=12, bsk_level=3, ksk_base_log=4, ksk_level=6, lwe_std=2.0**-40, glwe_std=2.0**-55)
        ctx = Context(rank)
        q = C.quantize_input(circ, calib[:1].numpy())[0]
        single = CircuitExecutor(circ, (tlu, bit), ctx=ctx, input_std=2.0**-50)
        single.keygen(seed=3)
        cts = single.encrypt(q, enc_seed=4)
        want = ctx.to_host_u64(single.run(cts))
        multi = CircuitExecutor(circ, (tlu, bit), ctx=ctx, rank=rank, world_size=world, process_group=dist.group.WORLD, input_std=2.0**-50)
        multi.use_keys(single.keys)
        got = ctx.to_host_u64(multi.run(cts))
        ok = bool(np.array_equal(got, want))
        # default seeds come from the OS CSPRNG: rank 0 draws, every rank must end up with the same keys and the same input ciphertexts
        multi.keygen()
        import hashlib
        digest = hashlib.sha256(multi.keys.get_secret(-1).tobytes() + ctx.to_host_u64(multi.encrypt(q)).tobytes()).hexdigest()
        all_d = [None] * world
        dist.all_gather_object(all_d, digest)
        fresh = CircuitExecutor(circ, (tlu, bit), ctx=ctx, input_std=2.0**-50)
        fresh.keygen()
        ok = ok and len(set(all_d)) == 1 and not np.array_equal(fresh.keys.get_secret(-1), multi.keys.get_secret(-1))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_multi_gpu_matches_single_gpu():
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 8)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29700 + os.getpid() % 1000, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
