"""Run the reference's UNMODIFIED homomorphic_eval.py on top of this backend (the drop-in check of SURVEY 8(b)).

    python tools/run_reference_eval.py [--reference DIR] [--workdir DIR] [--synthetic-cifar N] -- <homomorphic_eval.py args>

e.g. the reference's DCT ResNet-20 configuration (run_homomorphic_eval.sh:45-49) on one image, encrypted:

    python tools/run_reference_eval.py -- --dataset cifar10 --model ResNet20qat --dct_status --channels 24 \
        --filter_size 4 --image_size_dct 16 --bit_width 4 --fhe_mode execute --calib_batch_size 100 \
        --test_batch_size 1 --test_subset 1 --rounding_threshold_bits 6 --n_bits 5 --p_error 0.01

What this wrapper does — and nothing else:
  * puts `dct-cryptonets_b200/` first on sys.path, so `concrete.ml.torch.compile`, `concrete.fhe` resolve to this backend's
    mirror of the Concrete-ML calls (homomorphic_eval.py:22-23);
  * appends `dct-cryptonets_b200/compat/` LAST on sys.path: import shims for packages the script imports but this image
    lacks (brevitas, turbojpeg, jpeg2dct, torchinfo, matplotlib, seaborn) — real installations win;
  * CIFAR-10 cannot be downloaded (no network): with --synthetic-cifar N it writes N random train and N random test images
    in torchvision's on-disk format under <workdir>/cifardataset and turns torchvision's md5 check off;
  * runs the script with runpy from <workdir> (it writes mlir.txt into the current directory).
The script itself, its models/, data/ and utils are loaded from --reference untouched.
fhe_circuit.keygen() needs a B200: on a machine without CUDA the script stops there with this backend's error (no CPU fallback).
"""
import argparse
import os
import pickle
import runpy
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_synthetic_cifar(root: str, n: int, seed: int = 0) -> None:
    import numpy as np
    base = os.path.join(root, "cifar-10-batches-py")
    os.makedirs(base, exist_ok=True)
    rng = np.random.default_rng(seed)
    names = ["airplane", "automobile", "bird", "cat", "deer", "dog", "frog", "horse", "ship", "truck"]

    def batch(count):
        # smooth random fields (so the DCT coefficients look like a picture's, not like white noise)
        low = rng.integers(0, 256, size=(count, 3, 8, 8)).astype(np.float32)
        img = np.repeat(np.repeat(low, 4, axis=2), 4, axis=3) + rng.normal(0, 12, size=(count, 3, 32, 32))
        data = np.clip(img, 0, 255).astype(np.uint8).reshape(count, 3072)
        return {"data": data, "labels": [int(v) for v in rng.integers(0, 10, size=count)]}

    per = max(1, n // 5)
    for i in range(1, 6):
        with open(os.path.join(base, f"data_batch_{i}"), "wb") as f:
            pickle.dump(batch(per), f)
    with open(os.path.join(base, "test_batch"), "wb") as f:
        pickle.dump(batch(max(2, n)), f)
    with open(os.path.join(base, "batches.meta"), "wb") as f:
        pickle.dump({"label_names": names, "num_cases_per_batch": per, "num_vis": 3072}, f)


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default="/root/reference/dct-cryptonets", help="directory holding the unmodified homomorphic_eval.py")
    ap.add_argument("--workdir", default=None, help="directory to run in (default: a fresh temporary directory)")
    ap.add_argument("--synthetic-cifar", type=int, default=200, metavar="N", help="write N synthetic CIFAR-10 images (0: use what is in workdir)")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    args.reference = os.path.abspath(args.reference)                   # the script runs from <workdir>
    script = os.path.join(args.reference, "homomorphic_eval.py")
    if not os.path.isfile(script):
        raise SystemExit(f"{script} not found (the reference tree is only present in the build container)")
    rest = args.script_args[1:] if args.script_args[:1] == ["--"] else args.script_args

    pkg = os.path.join(ROOT, "dct-cryptonets_b200")
    sys.path.insert(0, args.reference)
    sys.path.insert(0, pkg)
    sys.path.append(os.path.join(pkg, "compat"))
    os.environ.setdefault("BREVITAS_IGNORE_MISSING_KEYS", "1")            # run_homomorphic_eval.sh:9

    workdir = args.workdir or tempfile.mkdtemp(prefix="tfx_ref_eval_")
    os.makedirs(workdir, exist_ok=True)
    if args.synthetic_cifar > 0:
        write_synthetic_cifar(os.path.join(workdir, "cifardataset"), args.synthetic_cifar)
        import torchvision.datasets.cifar as tv_cifar
        tv_cifar.check_integrity = lambda *a, **k: True                    # synthetic files cannot match the published md5 sums
    os.chdir(workdir)
    sys.argv = [script] + rest
    print(f"[run_reference_eval] {script} {' '.join(rest)}   (cwd {workdir})", flush=True)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
