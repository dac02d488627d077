"""What one rank of a W-GPU run costs, measured on ONE GPU: the executor runs rank 0's share of every lookup layer of the
headline circuit (DCT-ResNet-20) and the all-gather is replaced by a local tile of the rank's own block (timing only — the
outputs are not a valid inference).  Used to size the concurrent wave-sized chains of the lookup layers (executor.py) without paying for W GPUs.

    python tools/shard_probe.py --world 8 [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dct-cryptonets_b200"))

import bench                                            # noqa: E402
from tfx_b200 import circuit as C                       # noqa: E402
from tfx_b200.binding import Context                    # noqa: E402
from tfx_b200.executor import CircuitExecutor           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", default="8")
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--chains", default="1,2,4", help="max concurrent wave-sized chains per lookup layer (executor.max_chains)")
    ap.add_argument("--json", default=None)
    ap.add_argument("--breakdown", action="store_true", help="per-kernel-class CUDA-event sums of one more run")
    ap.add_argument("--layers", action="store_true", help="per-layer seconds against 1/world of the single-GPU layer time")
    ap.add_argument("--timeline", default=None, help="write (class, units, start ms, end ms) of every launch of one more run to this JSON file")
    args = ap.parse_args()
    _, circ, params, info, image = bench.build_circuit_and_params()
    ctx = Context(0)
    base = CircuitExecutor(circ, params, ctx=ctx)
    base.keygen(seed=1)
    cts = base.encrypt(C.quantize_input(circ, image)[0], enc_seed=2)
    rows = []
    full_layers = None
    if args.layers:
        from tfx_b200.executor import RunStats
        st = RunStats()
        base.run(cts, st, time_layers=True)
        full_layers = st.layer_seconds
    for world in [int(w) for w in args.world.split(",")]:
        for split in [int(c) for c in args.chains.split(",")]:
            ex = CircuitExecutor(circ, params, ctx=ctx, rank=0, world_size=world)
            ex.use_keys(base.keys)
            ex.max_chains = split

            def fake_gather(local, Cc, per, hw):
                reps = -(-Cc * hw // max(1, local.shape[0]))
                return local.repeat(reps, 1)[: Cc * hw]
            ex._gather = fake_gather
            ex.run(cts)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ex.run(cts); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 1e3)
            row = {"world": world, "max_chains": split, "rank0_seconds": min(ts)}
            if args.breakdown:
                from tfx_b200.executor import RunStats
                st = RunStats()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ex.run(cts, st, profile_kernels=True); e1.record()
                torch.cuda.synchronize()
                ks = st.kernel_seconds()
                row["profiled_seconds"] = e0.elapsed_time(e1) / 1e3
                row["classes"] = {k: [round(v[0], 4), v[1], v[2]] for k, v in ks.items()}
                row["class_sum"] = sum(v[0] for v in ks.values())
            if args.timeline:
                from tfx_b200.executor import RunStats
                st = RunStats()
                ref = torch.cuda.Event(enable_timing=True); ref.record()
                ex.run(cts, st, profile_kernels=True)
                torch.cuda.synchronize()
                tl = [(cls, int(u), ref.elapsed_time(a), ref.elapsed_time(b)) for cls, a, b, u in st.kernel_events]
                json.dump(tl, open(f"{args.timeline}.w{world}.c{split}.json", "w"))
            if args.layers:
                st = RunStats()
                ex.run(cts, st, time_layers=True)
                loss = sorted(((b - a / world, n, a / world, b) for (n, a), (_, b) in zip(full_layers, st.layer_seconds)), reverse=True)
                row["layer_loss_top"] = [(n, round(i * 1e3, 1), round(b * 1e3, 1)) for l, n, i, b in loss[:12]]
                row["layer_loss_total"] = sum(l for l, *_ in loss)
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
