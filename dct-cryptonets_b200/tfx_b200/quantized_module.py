"""QuantizedModule / FheCircuit: the objects the reference manipulates after compilation
(reference homomorphic_eval.py:301 graph.maximum_integer_bit_width(), :311 .mlir, :315 .keygen(), :70 .forward)."""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from . import circuit as C
from .params import pick_parameters


class _Graph:
    def __init__(self, circ: C.Circuit):
        self._circ = circ

    def maximum_integer_bit_width(self) -> int:
        return self._circ.maximum_integer_bit_width()


class FheCircuit:
    """compiled circuit + parameters; the GPU executor is created on keygen()"""

    def __init__(self, circ: C.Circuit, params, info: dict):
        self.circuit = circ
        self.params = params
        self.parameter_info = info
        self.graph = _Graph(circ)
        self.executor = None
        # None = drawn from the OS CSPRNG (keys at keygen, encryption masks per key set with a running PRF index);
        # fixed values only through keygen(seed=..., encryption_seed=...) for tests and benchmarks
        self.key_seed = None
        self.encryption_seed = None
        self.last_run_stats = None
        self.last_run_events = None        # CUDA events around executor.run (device time of the server-side step)
        self.last_output = None            # output ciphertexts of the last run (device tensor)
        self.profile_kernels = False       # bench.py: per-kernel-class CUDA events in last_run_stats
        self._dist = None

    @property
    def mlir(self) -> str:
        return self.circuit.to_text()

    @property
    def statistics(self) -> dict:
        s = dict(self.circuit.pbs_count())
        s["macs"] = self.circuit.macs()
        s["params"] = {"tlu": self.params[0], "bit": self.params[1]}
        return s

    def configure_distributed(self, rank: int, world_size: int, process_group=None):
        self._dist = (rank, world_size, process_group)

    def _ensure_executor(self):
        if self.executor is None:
            from .executor import CircuitExecutor      # needs CUDA; there is no CPU execution path
            rank, world, pg = self._dist if self._dist is not None else (0, 1, None)
            self.executor = CircuitExecutor(self.circuit, self.params, rank=rank, world_size=world, process_group=pg)
        return self.executor

    def keygen(self, force: bool = False, seed: Optional[int] = None, encryption_seed: Optional[int] = None):
        ex = self._ensure_executor()
        if seed is not None:
            self.key_seed = seed
        if encryption_seed is not None:
            self.encryption_seed = encryption_seed
        if ex.keys is None or force:
            ex.keygen(self.key_seed)

    def encrypt(self, q_x: np.ndarray):
        self.keygen()
        return self.executor.encrypt(q_x, self.encryption_seed)

    def run(self, cts):
        from .executor import RunStats
        self.last_run_stats = RunStats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = self.executor.run(cts, self.last_run_stats, profile_kernels=self.profile_kernels)
        e1.record()
        self.last_run_events, self.last_output = (e0, e1), out
        return out

    def decrypt(self, cts) -> np.ndarray:
        return self.executor.decrypt(cts)

    def encrypt_run_decrypt(self, q_x: np.ndarray) -> np.ndarray:
        return self.decrypt(self.run(self.encrypt(q_x)))

    def simulate(self, q_x: np.ndarray) -> np.ndarray:
        return C.evaluate_clear(self.circuit, q_x[None])[0]


class QuantizedModule:
    def __init__(self, fhe_circuit: FheCircuit, torch_model: torch.nn.Module):
        self.fhe_circuit = fhe_circuit
        self._model = torch_model

    @classmethod
    def compile(cls, torch_model, inputset: torch.Tensor, n_bits: int, rounding_threshold_bits: int, p_error: float,
                configuration=None, verbose: bool = False, params=None, rounding_method: str = "exact",
                per_channel_offsets: bool = True, per_channel_widths: Optional[bool] = None,
                fuse_residual: bool = False) -> "QuantizedModule":
        t0 = time.time()
        circ = C.build_circuit(torch_model, inputset, n_bits=n_bits, rounding_threshold_bits=rounding_threshold_bits, p_error=p_error,
                               rounding_method=rounding_method, per_channel_offsets=per_channel_offsets,
                               per_channel_widths=per_channel_widths, fuse_residual=fuse_residual)
        if params is None:
            tlu, bit, info = pick_parameters(circ.noise_spec())
        else:
            tlu, bit = params
            info = {"forced": True}
        info["compile_seconds"] = time.time() - t0
        if verbose:
            cnt = circ.pbs_count()
            print(f"[tfx_b200] circuit: {len(circ.ops)} ops, max width {circ.maximum_integer_bit_width()} bits, "
                  f"{cnt['tlu']} table lookups + {cnt['bit']} bit extractions = {cnt['total']} PBS / image")
            print(f"[tfx_b200] tlu set: {tlu}")
            print(f"[tfx_b200] bit set: {bit}")
        return cls(FheCircuit(circ, (tlu, bit), info), torch_model)

    # ---- the reference's hot call (homomorphic_eval.py:70) ---------------------------------------------
    def quantize_input(self, x: np.ndarray) -> np.ndarray:
        return C.quantize_input(self.fhe_circuit.circuit, x)

    def dequantize_output(self, q: np.ndarray) -> np.ndarray:
        return C.dequantize_output(self.fhe_circuit.circuit, q)

    def forward(self, x, fhe: str = "disable") -> np.ndarray:
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        x = np.asarray(x)
        if fhe not in ("disable", "simulate", "execute"):
            raise ValueError(f"fhe must be 'disable', 'simulate' or 'execute', got {fhe!r}")
        q_x = self.quantize_input(x)
        if fhe == "disable":
            q_y = C.evaluate_clear(self.fhe_circuit.circuit, q_x)
        elif fhe == "simulate":
            # clear integers + the modelled PBS noise of the picked parameter sets (p_error), like Concrete's simulation
            fc = self.fhe_circuit
            nm = C.NoiseModel.from_params(fc.params[0], fc.params[1], fc.params[0].glwe_std)
            self._sim_calls = getattr(self, "_sim_calls", 0) + 1
            q_y = C.evaluate_clear(fc.circuit, q_x, noise=nm, rng=np.random.default_rng(self._sim_calls))
        else:
            q_y = np.stack([self.fhe_circuit.encrypt_run_decrypt(q_x[i]) for i in range(q_x.shape[0])])
            q_y = q_y.reshape(q_x.shape[0], *self.fhe_circuit.circuit.output_shape)
        return self.dequantize_output(q_y)

    __call__ = forward
