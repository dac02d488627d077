"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/tfx.h declares,
fails loudly (no CPU fallback) when asked to compute, and its FFT tables equal the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tfx_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = binding.load_library()
    header = open(os.path.join(ROOT, "include", "tfx.h")).read()
    declared = set(re.findall(r"\b(tfx_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tfx.h but not exported"
    assert declared == set(binding.SYMBOLS), "binding.SYMBOLS out of sync with include/tfx.h"
    assert b"sm_100a" in lib.tfx_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = binding.load_library()
    h = C.c_void_p()
    rc = lib.tfx_ctx_create(0, None, 0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.tfx_last_error()
    with pytest.raises(binding.TfxError):
        binding.Context(0)


def test_fft_tables_equal_oracle(oracle):
    for N in (512, 1024, 2048, 4096, 8192):
        tw = binding.fft_tables(N)
        assert np.array_equal(tw, oracle.fft_tables(N))
        # first pass: one radix-8 node with rho = exp(i*pi/16): its 4th power is exactly (1+i)/sqrt(2);
        # N = 2048 (M = 1024, plan {2,2,3,3}) starts with a radix-4 node, rho = exp(i*pi/8): its 2nd power
        q = 1 if N == 2048 else 3
        assert tw[q, 0] == tw[q, 1] == np.float64(0.5) ** 0.5
        assert np.allclose(tw[:-1, 0] ** 2 + tw[:-1, 1] ** 2, 1.0, atol=1e-15)


def test_pbs_supported_matrix():
    lib = binding.load_library()
    assert lib.tfx_pbs_supported(4096, 1) == 1 and lib.tfx_pbs_supported(2048, 2) == 1
    assert lib.tfx_pbs_supported(4096, 3) == 0 and lib.tfx_pbs_supported(100, 1) == 0


def test_reference_arm_prints_the_contract_line_with_all_host_threads(tmp_path):
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with impl / metric / value / e2e /
    cpu_baseline, and it must use every host core even when torch.distributed.run has exported OMP_NUM_THREADS=1 (round 1 ran
    single-threaded under torchrun and timed out)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "16"], capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "encrypted_image_latency" and line["unit"] == "s/image"
    assert line["higher_is_better"] is False and line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert line["cpu_baseline"]["cores"] == cores
    # the other ranks exit without work and without output
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", LOCAL_RANK="1"), cwd=str(tmp_path))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
