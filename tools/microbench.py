"""Batched PBS + keyswitch microbenchmark at the circuit's TFHE parameter sets (BASELINE.json configs[1]).

python tools/microbench.py [--batches 148,1184,...] [--sets tlu,bit] [--json out.json]
Times each kernel with CUDA events on the launching stream (3 warm-ups, inputs larger than L2 or L2 flushed).
Under torch.distributed.run (N ranks, one per GPU) a batch of B ciphertexts is sharded B/N per rank — keys replicated, no
data-path collective — and the reported rates are whole-job (B over the slowest rank's time).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dct-cryptonets_b200"))

from tfx_b200 import params as P                      # noqa: E402
from tfx_b200.binding import Context, KeySet, PbsParams   # noqa: E402

# the sets the picker returns for DCT-ResNet-20 / 24x16^2 / n_bits 5 / rounding 6 / p_error 0.01 (seed 0)
DEFAULT_SETS = {
    # big LWE key of 2048 bits (the picker's choice since the big dimension is searched, tfx_b200/params.py)
    "tlu": PbsParams(n=768, k=1, N=2048, bsk_base_log=15, bsk_level=2, ksk_base_log=2, ksk_level=8,
                     lwe_std=P.min_noise_std(768), glwe_std=P.min_noise_std(2048)),
    "bit": PbsParams(n=492, k=2, N=1024, bsk_base_log=23, bsk_level=1, ksk_base_log=2, ksk_level=5,
                     lwe_std=P.min_noise_std(492), glwe_std=P.min_noise_std(2048)),
    # the sets of the earlier 4096-bit big key (profiles/r01_microbench_v1..v10): --sets tlu4096,bit4096
    "tlu4096": PbsParams(n=752, k=1, N=4096, bsk_base_log=16, bsk_level=2, ksk_base_log=2, ksk_level=7,
                         lwe_std=P.min_noise_std(752), glwe_std=P.min_noise_std(4096)),
    # bit-extraction set whose GLWE key is the first k*N = 2048 bits of the 4096-bit big key
    "bit4096": PbsParams(n=516, k=2, N=1024, bsk_base_log=23, bsk_level=1, ksk_base_log=2, ksk_level=5,
                         lwe_std=P.min_noise_std(516), glwe_std=P.min_noise_std(2048)),
    "bit_k1": PbsParams(n=516, k=1, N=2048, bsk_base_log=23, bsk_level=1, ksk_base_log=2, ksk_level=5,
                        lwe_std=P.min_noise_std(516), glwe_std=P.min_noise_std(2048)),
    "bit_full": PbsParams(n=524, k=2, N=2048, bsk_base_log=24, bsk_level=1, ksk_base_log=2, ksk_level=5,
                          lwe_std=P.min_noise_std(524), glwe_std=P.min_noise_std(4096)),
}


def time_fn(fn, warmup=3, iters=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 1e3)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="148,1184,4096")
    ap.add_argument("--sets", default="tlu,bit")
    ap.add_argument("--json", default=None)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-ks", action="store_true")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    names = args.sets.split(",")
    sets = [DEFAULT_SETS[n] for n in names]
    keys = KeySet.generate(ctx, sets, 1)
    dfma = ctx.probe_rate(0)
    imac = ctx.probe_rate(1)
    if rank == 0:
        print(f"measured DFMA peak {dfma / 1e12:.2f} TFLOP/s, u64 MAC peak {imac / 1e12:.2f} TMAC/s", flush=True)

    def slowest(t):
        if world == 1:
            return t
        v = torch.tensor([t], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=ctx.device)     # 256 MB > L2
    rows = []
    for Btot in [int(b) for b in args.batches.split(",")]:
        B = Btot // world + (1 if rank < Btot % world else 0)          # this rank's shard
        for sid, (name, p) in enumerate(zip(names, sets)):
            if B == 0:                                   # fewer ciphertexts than ranks: this rank idles
                t_pbs = slowest(0.0)
                if not args.no_ks:
                    slowest(0.0)
                continue
            g = torch.Generator(device="cuda"); g.manual_seed(Btot * 131 + rank)
            big = torch.randint(-2**62, 2**62, (B, keys.big_dim + 1), dtype=torch.int64, device=ctx.device, generator=g)
            small = torch.randint(-2**62, 2**62, (B, p.n + 1), dtype=torch.int64, device=ctx.device, generator=g)
            luts = torch.randint(-2**62, 2**62, (4, p.N), dtype=torch.int64, device=ctx.device, generator=g)
            idx = torch.zeros(B, dtype=torch.int32, device=ctx.device)
            out = ctx.empty_u64(B, keys.big_dim + 1)
            ks_out = ctx.empty_u64(B, p.n + 1)
            t_pbs = slowest(time_fn(lambda: keys.pbs(sid, small, luts, idx, out=out), args.warmup, args.iters, flush))
            row = {"set": name, "B": Btot, "n_gpus": world, "pbs_s": t_pbs, "pbs_per_s": Btot / t_pbs, "pbs_per_s_per_gpu": Btot / t_pbs / world,
                   "pbs_tflops": Btot * P.pbs_flops(p) / t_pbs / 1e12, "pbs_frac_dfma": Btot * P.pbs_flops(p) / t_pbs / dfma / world,
                   "latency_ms": t_pbs * 1e3,
                   "bsk_GBps_algorithmic": P.bsk_bytes(p) / t_pbs / 1e9,
                   "pbs_roofline": "min(key bytes at HBM rate, flops at DFMA peak)",
                   "pbs_roofline_frac": (max(P.bsk_bytes(p) / 6551.7e9, -(-Btot // world) * P.pbs_flops(p) / dfma)) / t_pbs}
            if not args.no_ks:
                t_ks = slowest(time_fn(lambda: keys.keyswitch(sid, big, out=ks_out), args.warmup, args.iters, flush))
                macs = P.ks_macs(p, keys.big_dim)
                row.update({"ks_s": t_ks, "ks_per_s": Btot / t_ks, "ks_tmacs": Btot * macs / t_ks / 1e12, "ks_frac_imac": Btot * macs / t_ks / imac / world})
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if args.json and rank == 0:
        with open(args.json, "w") as f:
            json.dump({"dfma_peak": dfma, "imac_peak": imac, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
