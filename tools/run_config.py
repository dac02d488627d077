"""Compile (and optionally run on cuda:0) the other BASELINE.json configurations: circuit statistics, picked TFHE
parameter sets, PBS counts; with --run, one encrypted inference and its latency.

    python tools/run_config.py --config 3|4|5 [--run]
config 3: ResNet-20 on raw RGB 3x32^2 (skip_single_downsample=False)
config 4: DCT-ResNet-18 on 24x16^2 (added stem key '64_24_16')
config 5: DCT-ResNet-18 on 64x56^2 (ImageNette-size DCT input), one image per GPU (replicas only)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dct-cryptonets_b200"))
from tfx_b200 import circuit as C, params as P                       # noqa: E402
from tfx_b200.resnet_dct import resnet18_dct, resnet20_dct           # noqa: E402


def build(config: int, calib_n: int):
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    if config == 1:
        model, shape = resnet20_dct(24, 16), (24, 16, 16)
    elif config == 3:
        model, shape = resnet20_dct(3, 32, skip_single_downsample=False), (3, 32, 32)
    elif config == 4:
        model, shape = resnet18_dct(24, 16), (24, 16, 16)
    elif config == 5:
        model, shape = resnet18_dct(64, 56), (64, 56, 56)
    else:
        raise SystemExit("config must be 1, 3, 4 or 5")
    calib = torch.randn(calib_n, *shape, generator=g)
    image = torch.randn(1, *shape, generator=g).numpy()
    return model.eval(), calib, image


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--calib", type=int, default=100)
    ap.add_argument("--run", action="store_true")
    args = ap.parse_args()
    model, calib, image = build(args.config, args.calib)
    t0 = time.time()
    circ = C.build_circuit(model, calib, n_bits=5, rounding_threshold_bits=6, p_error=0.01)
    t_compile = time.time() - t0
    tlu, bit, info = P.pick_parameters(circ.noise_spec())
    cnt = circ.pbs_count()
    out = {"config": args.config, "compile_s": t_compile, "ops": len(circ.ops), "lookup_layers": len(circ.lookups()),
           "max_bit_width": circ.maximum_integer_bit_width(), "conv_macs": circ.macs(), "pbs": cnt,
           "tlu_set": str(tlu), "bit_set": str(bit), "worst_margin": info["worst_margin"],
           "gflop_per_image": (cnt["tlu"] * P.pbs_flops(tlu) + cnt["bit"] * P.pbs_flops(bit)) / 1e9}
    if args.run:
        from tfx_b200.executor import CircuitExecutor, RunStats
        ex = CircuitExecutor(circ, (tlu, bit))
        out["keygen_s"] = ex.keygen(1)
        q = C.quantize_input(circ, image)[0]
        cts = ex.encrypt(q, 2)
        torch.cuda.synchronize()
        st = RunStats()
        t0 = time.time()
        res = ex.run(cts, st)
        torch.cuda.synchronize()
        out["latency_s"] = time.time() - t0
        dec = ex.decrypt(res)
        clear = C.evaluate_clear(circ, q[None])[0].reshape(-1)
        out["max_abs_dev_from_clear"] = int(np.abs(dec - clear).max())
        out["clear_span"] = int(clear.max() - clear.min())
        out["pbs_per_s"] = (st.pbs_tlu + st.pbs_bit) / out["latency_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
