"""What one rank of a W-GPU run costs, measured on ONE GPU: the executor runs rank 0's share of every lookup layer of the
headline circuit (DCT-ResNet-20) and the all-gather is replaced by a local tile of the rank's own block (timing only — the
outputs are not a valid inference).  Used to size the two-stream lookup layers (executor.py) without paying for W GPUs.

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
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    circ, params, info, image = bench.build_circuit_and_params()
    ctx = Context(0)
    base = CircuitExecutor(circ, params, ctx=ctx)
    base.keygen(seed=1)
    cts = base.encrypt(C.quantize_input(circ, image)[0], enc_seed=2)
    rows = []
    for world in [int(w) for w in args.world.split(",")]:
        for split in (False, True):
            ex = CircuitExecutor(circ, params, ctx=ctx, rank=0, world_size=world)
            ex.use_keys(base.keys)
            ex.split_streams = split

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
            row = {"world": world, "two_streams": split, "rank0_seconds": min(ts)}
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
