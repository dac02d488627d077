"""FHEModelDev / FHEModelClient / FHEModelServer with Concrete-ML's deployment call pattern (named in
BASELINE.json's north_star; the reference itself never calls them — SURVEY.md §0.2).

Bytes cross the client/server boundary: the client holds the secret keys, the server only evaluation keys
(keyswitch keys + Fourier bootstrapping keys) and ciphertexts.  Wire format (little endian):
    b"TFXB" u32 version | u64 json_len | json header | raw arrays in header order
"""
from __future__ import annotations

import hashlib
import io
import json
import os
import struct
import zipfile
from typing import Optional

import numpy as np
import torch

from tfx_b200.binding import Context, KeySet, PbsParams
from tfx_b200.executor import CircuitExecutor
from tfx_b200 import circuit as C

_MAGIC = b"TFXB"
_VERSION = 1


def _pack(header: dict, arrays) -> bytes:
    header = dict(header)
    header["arrays"] = [{"dtype": str(a.dtype), "shape": list(a.shape)} for a in arrays]
    hj = json.dumps(header).encode()
    out = io.BytesIO()
    out.write(_MAGIC + struct.pack("<IQ", _VERSION, len(hj)) + hj)
    for a in arrays:
        out.write(np.ascontiguousarray(a).tobytes())
    return out.getvalue()


def _unpack(blob: bytes):
    if blob[:4] != _MAGIC:
        raise ValueError("not a tfx_b200 blob")
    ver, hlen = struct.unpack("<IQ", blob[4:16])
    if ver != _VERSION:
        raise ValueError(f"unsupported blob version {ver}")
    header = json.loads(blob[16:16 + hlen])
    off = 16 + hlen
    arrays = []
    for spec in header["arrays"]:
        n = int(np.prod(spec["shape"])) * np.dtype(spec["dtype"]).itemsize
        arrays.append(np.frombuffer(blob, dtype=spec["dtype"], count=int(np.prod(spec["shape"])), offset=off).reshape(spec["shape"]))
        off += n
    return header, arrays


def _params_to_json(params):
    return [vars(p) for p in params]


def _params_from_json(lst):
    return [PbsParams(**d) for d in lst]


class FHEModelDev:
    """Saves what client and server need: the compiled circuit and its TFHE parameter sets."""

    def __init__(self, path_dir: str, model=None):
        self.path_dir, self.model = path_dir, model

    def save(self, via_mlir: bool = False):
        os.makedirs(self.path_dir, exist_ok=True)
        fc = self.model.fhe_circuit
        head, arrays = C.circuit_to_portable(fc.circuit)
        payload = _pack({"kind": "circuit", "circuit": head, "params": _params_to_json(fc.params)}, arrays)
        for name in ("client.zip", "server.zip"):
            path = os.path.join(self.path_dir, name)
            if os.path.exists(path):
                raise FileExistsError(f"{path} already exists")
            with zipfile.ZipFile(path, "w") as z:
                z.writestr("circuit.tfxb", payload)           # JSON header + raw arrays: loading a bundle executes no code
                z.writestr("versions.json", json.dumps({"tfx_b200": _VERSION}))


def _load_bundle(path_dir: str, name: str):
    with zipfile.ZipFile(os.path.join(path_dir, name)) as z:
        header, arrays = _unpack(z.read("circuit.tfxb"))
    if header.get("kind") != "circuit":
        raise ValueError("expected a circuit bundle")
    return C.circuit_from_portable(header["circuit"], arrays), _params_from_json(header["params"])


class FHEModelClient:
    def __init__(self, path_dir: str, key_dir: Optional[str] = None):
        self.path_dir, self.key_dir = path_dir, key_dir
        self.circuit, self.params = _load_bundle(path_dir, "client.zip")
        self._ex: Optional[CircuitExecutor] = None

    def _executor(self) -> CircuitExecutor:
        if self._ex is None:
            self._ex = CircuitExecutor(self.circuit, self.params)
        return self._ex

    def generate_private_and_evaluation_keys(self, force: bool = False):
        ex = self._executor()
        if ex.keys is None or force:
            ex.keygen()                        # fresh 16 bytes from the OS CSPRNG (executor.keygen)

    def get_serialized_evaluation_keys(self) -> bytes:
        self.generate_private_and_evaluation_keys()
        keys = self._executor().keys
        arrays = []
        for s in range(len(self.params)):
            arrays += [keys.get_ksk(s), keys.get_bsk_fourier(s)]
        return _pack({"kind": "evaluation_keys", "params": _params_to_json(self.params)}, arrays)

    def quantize_encrypt_serialize(self, x: np.ndarray) -> bytes:
        self.generate_private_and_evaluation_keys()
        ex = self._executor()
        q = C.quantize_input(self.circuit, np.asarray(x))
        assert q.shape[0] == 1, "one sample per call"
        # encryption randomness: its own os.urandom seed per key set (independent of the key seed) with a PRF index that
        # advances by the ciphertexts produced so far (executor.encrypt) — masks never repeat, not even across calls
        cts = ex.encrypt(q[0])
        return _pack({"kind": "ciphertexts", "width": self.circuit.input_width}, [ex.ctx.to_host_u64(cts)])

    def deserialize_decrypt(self, blob: bytes) -> np.ndarray:
        _, (cts,) = _unpack(blob)
        ex = self._executor()
        q = ex.decrypt(ex.ctx.to_device_u64(cts))
        return q.reshape(1, *self.circuit.output_shape)

    def deserialize_decrypt_dequantize(self, blob: bytes) -> np.ndarray:
        return C.dequantize_output(self.circuit, self.deserialize_decrypt(blob))


class FHEModelServer:
    def __init__(self, path_dir: str):
        self.path_dir = path_dir
        self.circuit = None
        self.params = None
        self._ex: Optional[CircuitExecutor] = None
        self._key_digest = None
        self.load()

    def load(self):
        self.circuit, self.params = _load_bundle(self.path_dir, "server.zip")

    def _install_keys(self, blob: bytes):
        digest = hashlib.sha256(blob).hexdigest()              # the whole blob: two key sets never share a digest
        if self._ex is not None and digest == self._key_digest:
            return
        header, arrays = _unpack(blob)
        if header.get("kind") != "evaluation_keys":
            raise ValueError("expected serialized evaluation keys")
        params = _params_from_json(header["params"])
        if self._ex is not None and list(self._ex.params) != list(params):
            self._ex = None                                    # another client's parameter sets: rebuild the executor
        if self._ex is None:
            self._ex = CircuitExecutor(self.circuit, params)
        ks = KeySet.empty(self._ex.ctx, params)
        for s in range(len(params)):
            ks.set_ksk(s, arrays[2 * s])
            ks.set_bsk_fourier(s, arrays[2 * s + 1])
        self._ex.use_keys(ks)
        self._key_digest = digest

    def run(self, serialized_encrypted_quantized_data: bytes, serialized_evaluation_keys: bytes) -> bytes:
        self._install_keys(serialized_evaluation_keys)
        header, (cts,) = _unpack(serialized_encrypted_quantized_data)
        ex = self._ex
        out = ex.run(ex.ctx.to_device_u64(cts))
        return _pack({"kind": "ciphertexts", "width": self.circuit.output_width}, [ex.ctx.to_host_u64(out)])
