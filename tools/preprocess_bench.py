"""Throughput of the DCT preprocessing row: the batched torch pipeline on the GPU against the numpy oracle
(the per-image CPU restatement of the reference's transform) on the host.

    python tools/preprocess_bench.py [--batch 4096] [--cpu-images 64] [--json out.json]
Inputs are synthetic CIFAR-sized RGB images (32x32), configuration = the headline one (24 channels, 16x16, 4x4 blocks).
GPU time: CUDA events, images resident on the device, 3 warm-ups; a second figure includes the H2D copy of the uint8 batch from
pinned memory and the D2H copy of the float32 result.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dct-cryptonets_b200"))
from tfx_b200.dct_preprocess import DctPreprocessor          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--cpu-images", type=int, default=64)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(args.batch, 32, 32, 3), dtype=np.uint8)
    dev = torch.device(args.device)
    pre = DctPreprocessor(16, 4, 24, device=dev)
    out = {"config": "24 channels, 16x16, 4x4 block DCT, 32x32 RGB inputs", "batch": args.batch, "device": str(dev)}
    if dev.type == "cuda":
        host = torch.from_numpy(imgs).pin_memory()
        x = host.to(dev)
        for _ in range(3):
            y = pre(x)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        res = torch.empty(args.batch, 24, 16, 16, dtype=torch.float32).pin_memory()
        e[0].record(); y = pre(x); e[1].record()
        e[2].record(); y2 = pre(host.to(dev, non_blocking=True)); res.copy_(y2, non_blocking=True); e[3].record()
        torch.cuda.synchronize()
        out["gpu_images_per_s_resident"] = args.batch / (e[0].elapsed_time(e[1]) / 1e3)
        out["gpu_images_per_s_host_to_host"] = args.batch / (e[2].elapsed_time(e[3]) / 1e3)
        got = y[: args.cpu_images].cpu().numpy()
    else:
        t0 = time.time(); y = pre(imgs); dt = time.time() - t0
        out["torch_cpu_images_per_s"] = args.batch / dt
        got = y[: args.cpu_images].numpy()
    from oracle import dct_oracle as DO                        # checker + CPU baseline only
    t0 = time.time()
    ref = np.stack([DO.preprocess(imgs[i], 16, 4, 24) for i in range(args.cpu_images)])
    out["cpu_oracle_images_per_s"] = args.cpu_images / (time.time() - t0)
    out["cpu_oracle_note"] = "numpy restatement of the reference's per-image transform, one host thread"
    tol = 2 * np.spacing(np.maximum(np.abs(ref), np.float32(1e-3)))
    out["max_ulp_violations"] = int((np.abs(got - ref) > tol).sum())
    print(json.dumps(out))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
