#!/usr/bin/env python
"""bench.py — encrypted-image latency of DCT-ResNet-20 (24x16^2, CIFAR-10 shape) on N B200s, plus PBS/s/GPU.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W    (CPU oracle on the host cores, bounded sample)

A step is one server-side run of the compiled circuit on one synthetic encrypted image (reference timed region:
homomorphic_eval.py:350-363, minus the clear-text pre/post-processing).  `value` is the latency with the input
ciphertexts already resident in HBM; `e2e` is the same step measured through the host-facing call with the
ciphertexts in pinned host memory (H2D of the 6144 input ciphertexts and D2H of the 64 output ciphertexts inside
the timed region).  N > 1 partitions every table-lookup layer over the ranks (strong scaling of one image).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dct-cryptonets_b200"))
sys.path.insert(0, ROOT)

METRIC = "encrypted_image_latency"
UNIT = "s/image"
# BASELINE.json configs (configs[1], the kernel sweep, is tools/microbench.py).  Deviations from the reference's stem
# tables for 3 and 4 are SURVEY.md 3.4's; 5 runs one image per GPU (replicas only).
CONFIGS = {
    1: ("DCT-ResNet-20, one synthetic 24x16x16 DCT-domain image, n_bits=5, rounding_threshold_bits=6, p_error=0.01",
        lambda R: R.resnet20_dct(24, 16), (24, 16, 16), 6),
    3: ("ResNet-20 on raw RGB, one synthetic 3x32x32 image (skip_single_downsample=False), n_bits=5, rounding_threshold_bits=6, p_error=0.01",
        lambda R: R.resnet20_dct(3, 32, skip_single_downsample=False), (3, 32, 32), 6),
    4: ("DCT-ResNet-18, one synthetic 24x16x16 DCT-domain image (stem key '64_24_16'), n_bits=5, rounding_threshold_bits=6, p_error=0.01",
        lambda R: R.resnet18_dct(24, 16), (24, 16, 16), 6),
    5: ("DCT-ResNet-18, ImageNette-size 64x56x56 DCT-domain input, one synthetic image per GPU (replicas), n_bits=5, "
        "rounding_threshold_bits=6, p_error=0.01",
        lambda R: R.resnet18_dct(64, 56), (64, 56, 56), 6),
}
WORKLOAD = CONFIGS[1][0]


def build_circuit_and_params(config: int = 1, image_seed: int = 0):
    import torch
    from tfx_b200 import circuit as C, params as P, resnet_dct as R
    workload, make, shape, t_bits = CONFIGS[config]
    torch.manual_seed(0)
    model = make(R).eval()
    g = torch.Generator().manual_seed(0)
    calib = torch.randn(100, *shape, generator=g)
    circ = C.build_circuit(model, calib, n_bits=5, rounding_threshold_bits=t_bits, p_error=0.01)
    tlu, bit, info = P.pick_parameters(circ.noise_spec())
    image = torch.randn(1, *shape, generator=g).numpy()
    if image_seed:                                       # replicas: every rank its own image
        image = torch.randn(1, *shape, generator=torch.Generator().manual_seed(1000 + image_seed)).numpy()
    return model, circ, (tlu, bit), info, image


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) on a bounded sample
# ------------------------------------------------------------------------------------------------------------
_ORACLE_KEYS = None


def cpu_sample(circ, params, sample_cts: int = 32):
    """Times the oracle (C port, OpenMP) on `sample_cts` ciphertexts of the two PBS steps that make up >99 % of the
    image (bit-extraction step = keyswitch + PBS on the `bit` set; table step = keyswitch + PBS on the `tlu` set),
    then scales by the circuit's step counts.  Returns (extrapolated s/image, description, threads)."""
    global _ORACLE_KEYS
    from oracle import oracle as O, circuit_oracle as CO
    O.build()
    # all the host cores this process may use, set explicitly: torchrun exports OMP_NUM_THREADS=1 for nproc-per-node > 1
    O.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    tlu, bit = params
    if _ORACLE_KEYS is None:
        _ORACLE_KEYS = CO.OracleKeys(params, 1)
    keys = _ORACLE_KEYS
    rng = np.random.default_rng(0)
    big_dim = max(tlu.big_dim, bit.big_dim)
    acc = rng.integers(0, 2**64, size=(sample_cts, big_dim + 1), dtype=np.uint64)
    t0 = time.time()
    small = O.keyswitch(keys.ksk[1], acc, bit.ksk_base_log, bit.ksk_level, shift=12, body_offset=1 << 62)
    lut = np.full((1, bit.N), (-(1 << 50)) & (2**64 - 1), dtype=np.uint64)
    O.pbs(keys.bsk_f[1], bit.bsk_base_log, small, lut, np.zeros(sample_cts, np.uint32), mode=1, body_const=1 << 50, out=acc, big_dim=big_dim)
    t_bit = (time.time() - t0) / sample_cts
    t0 = time.time()
    small = O.keyswitch(keys.ksk[0], acc, tlu.ksk_base_log, tlu.ksk_level)
    lut = rng.integers(0, 2**64, size=(1, tlu.N), dtype=np.uint64)
    O.pbs(keys.bsk_f[0], tlu.bsk_base_log, small, lut, np.zeros(sample_cts, np.uint32))
    t_tlu = (time.time() - t0) / sample_cts
    cnt = circ.pbs_count()
    latency = cnt["bit"] * t_bit + cnt["tlu"] * t_tlu
    desc = (f"{sample_cts} ciphertexts through one bit-extraction step (KS+PBS, {t_bit * 1e3:.1f} ms/ct) and one table step "
            f"(KS+PBS, {t_tlu * 1e3:.1f} ms/ct) at the circuit's parameter sets; scaled by {cnt['bit']} + {cnt['tlu']} steps/image "
            f"(leveled convs, <1 % of the work, not included)")
    return latency, desc, O.num_threads()


def run_reference(args):
    """CPU arm: the oracle port on every host core, one bounded sample per step.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.time()
    _, circ, params, info, image = build_circuit_and_params(args.config)
    # size the per-step sample so that warmup + steps samples end within about two minutes on this host
    _, _, threads = cpu_sample(circ, params, sample_cts=8)                # also builds the oracle keys (untimed)
    t0 = time.time()
    cpu_sample(circ, params, sample_cts=max(8, threads))
    per_ct = (time.time() - t0) / max(8, threads)
    budget = 100.0 / max(1, args.warmup + args.steps)
    sample = int(min(args.cpu_sample, max(threads, budget / max(per_ct, 1e-6))))
    sample = max(threads, sample - sample % max(1, threads))
    lat = []
    desc = ""
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        v, desc, threads = cpu_sample(circ, params, sample_cts=sample)
        if it >= args.warmup:
            lat.append((v, time.time() - t0))
    value = statistics.mean(v for v, _ in lat)
    cnt = circ.pbs_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": statistics.mean(w for _, w in lat) * 1e3, "higher_is_better": False,
        "scaling": "weak" if args.config == 5 else "strong", "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config][0], "pbs_per_image": cnt["total"],
                   "note": "value is extrapolated from the bounded sample of each step; host cores only, no GPU",
                   "host_cores": threads, "wall_s": None},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pbs_per_sec": cnt["total"] / value,
        "published_context": "reference README.md:84 reports 565 s on 96 CPU cores with Concrete (not runnable here; parity unpinned)",
    }
    line["config"]["wall_s"] = time.time() - t_start
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import hashlib
    import torch
    import torch.distributed as dist
    from tfx_b200 import circuit as C, params as P
    from tfx_b200.binding import Context, launch_count
    from tfx_b200.quantized_module import FheCircuit, QuantizedModule

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    replicas = args.config == 5 and world > 1            # one image per GPU, no data-path collective

    model, circ, params, info, image = build_circuit_and_params(args.config, image_seed=rank if replicas else 0)
    tlu, bit = params
    # the reference-facing object: q_module.forward(numpy, fhe='execute') -> numpy (homomorphic_eval.py:70)
    qm = QuantizedModule(FheCircuit(circ, params, info), model)
    fc = qm.fhe_circuit
    if world > 1 and not replicas:
        fc.configure_distributed(rank, world, pg)
    fc.profile_kernels = True
    t0 = time.time()
    fc.keygen(seed=1, encryption_seed=2)                 # fixed seeds: every N reproduces the same ciphertext words (output_sha)
    torch.cuda.synchronize()
    t_keygen = time.time() - t0
    ex = fc.executor
    ctx = ex.ctx
    q_in = C.quantize_input(circ, image)[0]
    n_in, n_out = int(q_in.size), int(np.prod(circ.output_shape))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        t0 = time.perf_counter()
        y = qm.forward(image, fhe="execute")             # quantise, H2D, encrypt, run, decrypt, D2H, de-quantise
        return time.perf_counter() - t0, fc.last_run_events, fc.last_run_stats, y

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = launch_count()
    t0 = time.time()
    res = [step() for _ in range(args.steps)]
    barrier()
    wall = time.time() - t0
    launches = launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dev_s = sum(e0.elapsed_time(e1) for _, (e0, e1), _, _ in res) / 1e3 / args.steps
    e2e_s = sum(w for w, _, _, _ in res) / args.steps
    t = torch.tensor([wall / args.steps, dev_s, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_s, dev_s, e2e_s = [float(v) for v in t.cpu()]

    # correctness guard inside the bench: decrypted outputs must be in the clear evaluator's neighbourhood, and the hash of
    # the output ciphertext words must be the same at every N (same keys, same input ciphertexts, same circuit)
    out_words = ctx.to_host_u64(fc.last_output)
    output_sha = hashlib.sha256(out_words.tobytes()).hexdigest()[:16]
    dec = ex.decrypt(fc.last_output)
    clear = C.evaluate_clear(circ, q_in[None])[0].reshape(-1)
    span = max(1, int(clear.max() - clear.min()))
    max_dev = int(np.abs(dec - clear).max())

    if rank == 0:
        cnt = circ.pbs_count()
        ks: dict = {}
        for _, _, st, _ in res:                          # class -> (seconds, launches, units) over the timed steps, this rank
            for k_, (s_, n_, u_) in st.kernel_seconds().items():
                a_ = ks.get(k_, (0.0, 0, 0))
                ks[k_] = (a_[0] + s_, a_[1] + n_, a_[2] + u_)
        hbm_peak, peak_src = peaks()
        dfma = ctx.probe_rate(0)
        imac = ctx.probe_rate(1)
        dom = max(("pbs_bit", "pbs_tlu"), key=lambda k: ks.get(k, (0, 0, 0))[0])
        prof = {}
        ppath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(ppath):
            prof = json.load(open(ppath))
        total_kernel_s = sum(v[0] for v in ks.values()) or 1.0

        def pbs_roofline(cls, p_):
            s_, n_, u_ = ks.get(cls, (0.0, 0, 0))
            if not n_:
                return None
            f_ = u_ * P.pbs_flops(p_)
            return {"kernel": f"pbs_kernel<log2N={p_.N.bit_length() - 1},k={p_.k}> ({cls})", "bound": "fp64", "achieved": f_ / s_ / 1e12,
                    "peak": dfma / 1e12, "unit": "TFLOP/s", "frac": f_ / s_ / dfma,
                    "peak_source": "DFMA rate measured live by tfx_probe_rate (MEASURED_PEAKS.json has no FP64 figure)",
                    "flops_per_pbs": P.pbs_flops(p_), "pbs_per_launch": u_ / n_, "avg_launch_ms": s_ / n_ * 1e3,
                    "share_of_step": s_ / total_kernel_s, "pbs_per_s": u_ / s_,
                    "traffic": (prof.get(cls, {}).get("dram_bytes_per_unit") or 0) * u_ / n_ or None,
                    "traffic_source": prof.get("source"),
                    "hbm": {"achieved": n_ * P.bsk_bytes(p_) / s_ / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": n_ * P.bsk_bytes(p_) / s_ / 1e9 / hbm_peak, "peak_source": peak_src,
                            "note": "algorithmic bootstrapping-key bytes per launch (one pass over the key serves the whole batch)"}}
        sizes = [int(np.prod(op.shape)) for op in circ.lookups()]
        line = {
            "metric": METRIC, "value": dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_s * 1e3, "higher_is_better": False, "scaling": "weak" if replicas else "strong", "vs_baseline": None,
            "dtype": "u64+f64", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config][0], "config_id": args.config,
                       "parallelism": (f"{world} replicas, one image per GPU" if replicas else f"layer-partitioned x{world}"),
                       "pbs_per_image": cnt["total"], "pbs_tlu": cnt["tlu"], "pbs_bit": cnt["bit"], "conv_macs": circ.macs(),
                       "accumulator_layout": ("one centred offset and one width per channel" if circ_layout(circ) == 2 else
                                              "one centred offset per channel" if circ_layout(circ) == 1 else "tensor-wide offset and width (Concrete-like)")
                                             + " (tfx_b200/circuit.py, DESIGN.md 3)",
                       "tlu_set": str(tlu), "bit_set": str(bit),
                       "l2": (f"no flush: every lookup layer streams its ciphertext tensor ({min(x for x in sizes if x > 64) * ex.words * 8 / 1e6:.0f}"
                              f"-{max(sizes) * ex.words * 8 / 1e6:.0f} MB) several times between two uses of any buffer, and the "
                              f"{ex.keys.device_bytes / 1e6:.0f} MB of keys alternate per kernel; the working set per step exceeds L2 (126 MB)"),
                       "keygen_s": t_keygen, "key_bytes": ex.keys.device_bytes},
            "pbs_per_sec_per_gpu": cnt["total"] / dev_s / (1 if replicas else world),
            "images_per_step": world if replicas else 1,
            "e2e": {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": n_in * 8, "d2h_bytes_per_step": n_out * 8,
                    "call": "q_module.forward(numpy float32[1,C,H,W], fhe='execute') -> numpy (reference homomorphic_eval.py:70): host "
                            "quantisation, H2D of the plaintext words, encryption, circuit, decryption, D2H of the phases, de-quantisation; "
                            "host wall clock around the call, max over ranks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": pbs_roofline(dom, bit if dom == "pbs_bit" else tlu),
            "roofline_other_pbs_kernel": pbs_roofline("pbs_tlu" if dom == "pbs_bit" else "pbs_bit", tlu if dom == "pbs_bit" else bit),
            "roofline_whole_step": {"note": "all PBS flops of this rank's share of one image (SURVEY 8(d) formula) over the device time of the step; "
                                            "at N > 1 the chains of a layer run on several streams, so the per-class times above overlap and "
                                            "this is the per-rank figure to compare across N",
                                    "achieved": (cnt["tlu"] * P.pbs_flops(tlu) + cnt["bit"] * P.pbs_flops(bit)) / (1 if replicas else world) / dev_s / 1e12,
                                    "peak": dfma / 1e12, "unit": "TFLOP/s",
                                    "frac": (cnt["tlu"] * P.pbs_flops(tlu) + cnt["bit"] * P.pbs_flops(bit)) / (1 if replicas else world) / dev_s / dfma},
            "kernel_breakdown_s_per_step": {k: v[0] / args.steps for k, v in ks.items()},
            "kernel_breakdown_note": ("CUDA-event time per kernel class on the launching stream" +
                                      ("; with N > 1 the two halves of a layer run on two streams, so class times overlap and sum to more than the step" if world > 1 else "")),
            "imac_peak_tmacs": imac / 1e12,
            "output_sha": output_sha,
            "check": {"max_abs_deviation_from_clear": max_dev, "clear_output_span": span,
                      "note": "p_error=0.01 per PBS makes execute != clear by design; tests/ hold the bit-exact parity checks; "
                              "output_sha = sha256 of the output ciphertext words (identical at every N for the same config)"},
            "published_context": "reference README.md:84: 565 s on 96 CPU cores (Concrete CPU)",
        }
        if world == 1 and not args.no_cpu_baseline:
            v, desc, threads = cpu_sample(circ, params, sample_cts=args.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def circ_layout(circ) -> int:
    """2: per-channel offsets and widths, 1: per-channel offsets, 0: tensor-wide"""
    per_w = any(getattr(op, "chan_bits", None) is not None for op in circ.ops)
    per_o = any(isinstance(getattr(op, "offset", 0), np.ndarray) and np.unique(op.offset).size > 1 for op in circ.ops)
    return 2 if per_w else 1 if per_o else 0


# ------------------------------------------------------------------------------------------------------------
# DCT preprocessing row: python bench.py --preprocess [--batch N]
# ------------------------------------------------------------------------------------------------------------
def run_preprocess(args):
    """Throughput of the batched torch DCT pipeline on the GPU (CUDA events; images resident, and host to host from pinned
    memory) with the numpy oracle — the per-image CPU restatement of the reference's transform — as cpu_baseline and checker.
    Synthetic CIFAR-sized RGB images, the headline configuration (24 channels, 16x16, 4x4 blocks)."""
    import torch
    from tfx_b200.dct_preprocess import DctPreprocessor
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(args.batch, 32, 32, 3), dtype=np.uint8)
    dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    pre = DctPreprocessor(16, 4, 24, device=dev)
    out = {"metric": "dct_preprocess_throughput", "unit": "images/s", "config": {"workload": "24 channels, 16x16, 4x4 block DCT, 32x32 RGB inputs",
                                                                                 "batch": args.batch}, "device": str(dev), "data": "synthetic"}
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
        out["value"] = args.batch / (e[0].elapsed_time(e[1]) / 1e3)
        out["e2e"] = {"value": args.batch / (e[2].elapsed_time(e[3]) / 1e3), "unit": "images/s", "h2d_bytes_per_step": imgs.nbytes,
                      "d2h_bytes_per_step": res.numel() * 4}
        got = y[: args.cpu_images].cpu().numpy()
    else:
        t0 = time.time(); y = pre(imgs); dt = time.time() - t0
        out["value"] = args.batch / dt
        got = y[: args.cpu_images].numpy()
    from oracle import dct_oracle as DO                        # cpu_baseline leg + checker
    t0 = time.time()
    ref = np.stack([DO.preprocess(imgs[i], 16, 4, 24) for i in range(args.cpu_images)])
    out["cpu_baseline"] = {"value": args.cpu_images / (time.time() - t0), "unit": "images/s", "cores": 1, "kind": "port",
                           "sample": f"{args.cpu_images} images through oracle/dct_oracle.py (numpy restatement of the reference's per-image transform)"}
    tol = 2 * np.spacing(np.maximum(np.abs(ref), np.float32(1e-3)))
    out["check"] = {"values_beyond_2_ulp": int((np.abs(got - ref) > tol).sum()), "compared": int(ref.size)}
    print(json.dumps(out), flush=True)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tfx", choices=["tfx", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="BASELINE.json configuration (1 = headline)")
    ap.add_argument("--cpu-sample", type=int, default=256, help="ciphertexts per CPU sample step (about 10 s of host work on 16 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--preprocess", action="store_true", help="measure the DCT preprocessing row instead of the encrypted circuit")
    ap.add_argument("--batch", type=int, default=4096, help="--preprocess: images per batch")
    ap.add_argument("--cpu-images", type=int, default=64, help="--preprocess: images through the CPU oracle")
    args = ap.parse_args()
    if args.preprocess:
        run_preprocess(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
