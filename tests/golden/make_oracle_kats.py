"""Regenerates tests/golden/oracle_kats.json from the CPU oracle (run from the repo root).
These are regression vectors of OUR conventions, not reference outputs: the reference's arithmetic lives in
concrete-python 2.7.0, which cannot be installed here (parity unpinned)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

seed = 424242
params = dict(n=24, k=1, N=512, bsk_base_log=9, bsk_level=2, ksk_base_log=5, ksk_level=3, lwe_std=2.0**-30, glwe_std=2.0**-45)
big = O.gen_binary_key(seed, O.ST_BIGKEY, 0, 512)
small = O.gen_binary_key(seed, O.ST_SMALLKEY, 0, 24)
ksk = O.gen_ksk(big, small, 5, 3, 2.0**-30, seed)
bsk = O.gen_bsk(small, big, 1, 512, 9, 2, 2.0**-45, seed)
pts = [int(i) << 59 for i in range(6)]
cts = O.lwe_encrypt(big, 2.0**-40, np.array(pts, dtype=np.uint64), seed + 1)
sm = O.keyswitch(ksk, cts, 5, 3)
lut = (np.arange(512, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))[None]
out = O.pbs(O.bsk_to_fourier(bsk), 9, sm, lut, np.zeros(6, np.uint32))
dec_in = [0, 2**64 - 1, 2**63, 0x0123456789ABCDEF, 0xFEDCBA9876543210]
kats = {
    "seed": seed, "params": params,
    "prf": [int(v) for v in O.prf_fill(seed, 5, 3, 4)],
    "decompose_in": dec_in, "decompose_out": [int(v) for v in O.decompose(np.array(dec_in, dtype=np.uint64), 7, 3).reshape(-1)],
    "big_key_weight": int(big.sum()), "small_key_weight": int(small.sum()),
    "ksk_xor": int(np.bitwise_xor.reduce(ksk.reshape(-1))), "bsk_xor": int(np.bitwise_xor.reduce(bsk.reshape(-1))),
    "plaintexts": pts, "cts_xor": int(np.bitwise_xor.reduce(cts.reshape(-1))),
    "ks_bodies": [int(v) for v in sm[:, -1]], "pbs_bodies": [int(v) for v in out[:, -1]],
    "pbs_xor": int(np.bitwise_xor.reduce(out.reshape(-1))),
}
json.dump(kats, open(os.path.join(ROOT, "tests", "golden", "oracle_kats.json"), "w"), indent=1)
print("wrote", len(json.dumps(kats)), "bytes")
