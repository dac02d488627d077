"""Do PBS launches on different streams share the GPU?  (sizing of the concurrent chains in executor.py)
    python tools/stream_overlap_probe.py
"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dct-cryptonets_b200"))
sys.path.insert(0, ROOT)
from tfx_b200.binding import Context, KeySet
from tools.microbench import DEFAULT_SETS

ctx = Context(0)
sets = [DEFAULT_SETS["tlu"], DEFAULT_SETS["bit"]]
keys = KeySet.generate(ctx, sets, 1)
dev = ctx.device
side = []
for _ in range(3):
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        side.append((st, Context(0)))
main = torch.cuda.current_stream(dev)


def bufs(sid, B):
    p = sets[sid]
    g = torch.Generator(device="cuda"); g.manual_seed(B)
    small = torch.randint(-2**62, 2**62, (B, p.n + 1), dtype=torch.int64, device=dev, generator=g)
    big = torch.randint(-2**62, 2**62, (B, keys.big_dim + 1), dtype=torch.int64, device=dev, generator=g)
    luts = torch.randint(-2**62, 2**62, (1, p.N), dtype=torch.int64, device=dev, generator=g)
    return small, big, luts, torch.zeros(B, dtype=torch.int32, device=dev), ctx.empty_u64(B, keys.big_dim + 1)


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run_split(sid, sizes, steps, with_ks):
    """chains of `steps` x (KS +) PBS, chunk i of sizes[i] rows on its own stream, steps enqueued round-robin"""
    data = [bufs(sid, B) for B in sizes]
    lanes = [(main, ctx)] + side[: len(sizes) - 1]

    def fn():
        for st, _ in lanes[1:]:
            st.wait_stream(main)
        for _ in range(steps):
            for (st, c), (small, big, luts, idx, out) in zip(lanes, data):
                with torch.cuda.stream(st):
                    s_ = keys.keyswitch(sid, big, ctx=c) if with_ks else small
                    keys.pbs(sid, s_, luts, idx, out=out, ctx=c)
        for st, _ in lanes[1:]:
            main.wait_stream(st)
    return timeit(fn)


for sid, name, wave in ((1, "bit", 592), (0, "tlu", 296)):
    one = run_split(sid, [wave], 1, False)
    print(json.dumps({"set": name, "one_wave_ms": one}))
    for sizes in ([352, 240], [wave, wave, 352], [1536], [512, 512, 512], [wave * 2, 352]):
        for ks in (False, True):
            t = run_split(sid, sizes, 4, ks)
            print(json.dumps({"set": name, "chunks": sizes, "steps": 4, "keyswitch": ks, "ms": round(t, 2),
                              "ideal_ms": round(4 * one * sum(sizes) / wave, 2)}), flush=True)
