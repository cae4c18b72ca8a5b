#!/usr/bin/env python
"""Generate tests/golden/golden_c2.json: SHA-256 of the sorted key stream of BASELINE.json configs[1]
(uint64 key + uint64 payload, 1e9 uniform records, ascending) as produced by the COMPILED REFERENCE
(oracle/_ref = /root/reference/radixSort.hpp, unmodified), plus smaller sizes of the same generator.

    make -C oracle ref && python tests/golden/make_golden_c2.py        (needs ~40 GB of RAM, ~3 min)

Keys are mix64(seed + i) (tests/oracle_lib.py: the counter-based generator the GPU tests reproduce on the
device), payload = i.  The reference is not stable, so only the key stream is hashed; payloads are
checked by the test through keys0[payload] == sorted keys.
"""
from __future__ import annotations

import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as O  # noqa: E402

SEED = 12345
SIZES = [1 << 20, 1 << 25, 1_000_000_000]


def main():
    assert O.ref_available(), "the compiled reference (oracle/_ref) is needed"
    out = {}
    for n in SIZES:
        keys = np.empty(n, np.uint64)
        step = 1 << 26
        for s in range(0, n, step):
            c = min(step, n - s)
            keys[s:s + c] = O.mix64_numpy(s, c, SEED)
        pay = np.arange(n, dtype=np.uint64)
        t0 = time.time()
        O.ref_sort_soa(keys, [pay], True)
        dt = time.time() - t0
        assert bool(np.all(keys[:-1] <= keys[1:]))
        h = hashlib.sha256()
        for s in range(0, n, step):
            h.update(keys[s:s + step].tobytes())
        # payload parity with the reference itself: every payload still points at its key
        probe = np.random.default_rng(1).integers(0, n, size=1 << 16)
        assert np.array_equal(O.mix64_numpy(0, 1, 0)[:0], np.empty(0, np.uint64))
        src = pay[probe]
        regen = np.array([int(O.mix64_numpy(int(i), 1, SEED)[0]) for i in src[:256]], dtype=np.uint64)
        assert np.array_equal(regen, keys[probe[:256]])
        out[str(n)] = {"seed": SEED, "sha256_sorted_keys": h.hexdigest(), "head": [int(v) for v in keys[:3]],
                       "tail": int(keys[-1]), "reference_seconds": round(dt, 2)}
        print(n, out[str(n)], flush=True)
        del keys, pay
    (HERE / "golden_c2.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
