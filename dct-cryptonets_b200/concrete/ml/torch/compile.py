"""compile_torch_model / compile_brevitas_qat_model with the signatures used at reference
homomorphic_eval.py:276-295.  Both build the integer circuit with tfx_b200.circuit and pick TFHE parameters for
p_error; QAT modules (QuantConv2d / QuantReLU / QuantIdentity) are read by duck typing (bit widths from the module)."""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch

from tfx_b200.quantized_module import QuantizedModule


def _rounding(rounding_threshold_bits):
    """int or {"n_bits": int, "method": "exact" | "approximate"} (reference homomorphic_eval.py:279-280, README.md:96-113)"""
    if rounding_threshold_bits is None:
        return 16, "exact"
    if isinstance(rounding_threshold_bits, dict):
        method = str(rounding_threshold_bits.get("method", "exact")).lower()
        if method not in ("exact", "approximate"):
            raise ValueError(f"unknown rounding method {method!r}")
        return int(rounding_threshold_bits["n_bits"]), method
    return int(rounding_threshold_bits), "exact"


def compile_torch_model(torch_model: torch.nn.Module, torch_inputset, n_bits: Union[int, dict] = 8,
                        rounding_threshold_bits=None, p_error: Optional[float] = None, configuration=None,
                        verbose: bool = False, **kwargs) -> QuantizedModule:
    if isinstance(n_bits, dict):
        n_bits = int(n_bits.get("op_inputs", n_bits.get("model_inputs", 8)))
    if isinstance(torch_inputset, np.ndarray):
        torch_inputset = torch.from_numpy(torch_inputset)
    bits, method = _rounding(rounding_threshold_bits)
    # backend-specific options (not part of Concrete-ML's signature): accumulator layout, see tfx_b200/circuit.py
    layout = {k: kwargs[k] for k in ("per_channel_offsets", "per_channel_widths", "fuse_residual") if k in kwargs}
    return QuantizedModule.compile(torch_model, torch_inputset, n_bits=int(n_bits),
                                   rounding_threshold_bits=bits, rounding_method=method,
                                   p_error=0.01 if p_error is None else float(p_error), configuration=configuration,
                                   verbose=verbose, **layout)


def compile_brevitas_qat_model(torch_model: torch.nn.Module, torch_inputset, n_bits: Union[int, dict, None] = None,
                               rounding_threshold_bits=None, p_error: Optional[float] = None, configuration=None,
                               verbose: bool = False, **kwargs) -> QuantizedModule:
    return compile_torch_model(torch_model, torch_inputset, n_bits=8 if n_bits is None else n_bits,
                               rounding_threshold_bits=rounding_threshold_bits, p_error=p_error,
                               configuration=configuration, verbose=verbose, **kwargs)
