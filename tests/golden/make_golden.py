#!/usr/bin/env python
"""Generate tests/golden/*.npz and golden_c1.json from the COMPILED REFERENCE.

Run in the build container (needs /root/reference and an AVX-512 host):

    make -C oracle ref oracle && python tests/golden/make_golden.py

Every output array in the fixtures is what /root/reference/radixSort.hpp
(simd_sort::radix_sort::sort, radixSort.hpp:1761-1783) returned for the stored input;
nothing here comes from this repository's own sort.  The fixtures travel to the GPU
box, the reference does not.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as O  # noqa: E402

SIZES = [1, 2, 10, 16, 17, 100, 1000]
DISTS = ["Uniform", "Gaussian", "Zero", "ZeroOne", "AlmostSorted", "ReverseSorted"]


def float_specials(dt):
    fi = np.finfo(dt)
    tiny = np.nextafter(dt.type(0), dt.type(1))  # smallest denormal
    vals = [0.0, -0.0, np.inf, -np.inf, tiny, -tiny, fi.max, -fi.max, fi.tiny, -fi.tiny, 1.0, -1.0, 0.5, -0.5,
            1.5, -1.5, 2.0, -2.0, 1e-3, -1e-3, 3.25, -3.25, 1e10, -1e10, 7.0, -7.0]
    a = np.array(vals * 2, dtype=dt)  # 52 elements: > 16, so the sign bit is split first (SURVEY 8a quirk i)
    rng = np.random.default_rng(7)
    rng.shuffle(a)
    return a


def main():
    assert O.ref_available(), "compiled reference not available (make -C oracle ref; AVX-512 host needed)"
    soa = {}
    for dt in O.KEY_DTYPES:
        dt = np.dtype(dt)
        for up in (True, False):
            for dist in DISTS:
                for n in SIZES:
                    keys = O.make_keys(dist, dt, n, seed=1000 + n)
                    if dt.kind == "f" and n <= 16:
                        keys = np.where(keys == 0, dt.type(0.0), keys)  # no -0.0 below the radix threshold
                    out = keys.copy()
                    O.ref_sort_soa(out, [], up)
                    tag = f"{dt.name}|{int(up)}|{dist}|{n}"
                    soa["in|" + tag] = keys
                    soa["out|" + tag] = out
            if dt.kind == "f":
                keys = float_specials(dt)
                out = keys.copy()
                O.ref_sort_soa(out, [], up)
                tag = f"{dt.name}|{int(up)}|Specials|{len(keys)}"
                soa["in|" + tag] = keys
                soa["out|" + tag] = out
    np.savez_compressed(HERE / "soa_keys.npz", **soa)

    # AoS: records = key + key-derived payload bytes (src/data.hpp:393-406), all power-of-two sizes.
    # A record is a pure function of its key, so only the key columns are stored; the test rebuilds
    # the records with O.make_records and compares whole records.
    aos = {}
    for dt in O.KEY_DTYPES:
        dt = np.dtype(dt)
        rb = dt.itemsize
        while rb <= 64:
            for up in (True, False):
                for dist, n in (("Uniform", 1000), ("Gaussian", 100), ("ZeroOne", 17), ("Uniform", 10)):
                    keys = O.make_keys(dist, dt, n, seed=2000 + n + rb)
                    rec = O.make_records(keys, rb)
                    out = rec.copy()
                    O.ref_sort_aos(out, dt, up)
                    out_keys = np.ascontiguousarray(out[:, : dt.itemsize]).reshape(-1).view(dt)
                    assert np.array_equal(out, O.make_records(out_keys, rb))
                    tag = f"{dt.name}|{rb}|{int(up)}|{dist}|{n}"
                    aos["inkeys|" + tag] = keys
                    aos["outkeys|" + tag] = out_keys
            rb *= 2
    np.savez_compressed(HERE / "aos_records.npz", **aos)

    # config 1 (BASELINE.json configs[0]): Data<uint32_t,uint32_t>(1e6, Uniform, 42), ascending
    c1 = {}
    for n in (100_000, 1_000_000):
        k, p = O.c1_input(n, 42)
        first_keys, first_pay = k[:4].tolist(), p[:2].tolist()
        O.ref_sort_soa(k, [p], True)
        c1[str(n)] = {
            "input_first_keys": first_keys, "input_first_payloads": first_pay,
            "sorted_head": k[:3].tolist(), "sorted_tail": int(k[-1]),
            "sha256_sorted_keys": hashlib.sha256(k.tobytes()).hexdigest(),
            "sha256_sorted_payloads": hashlib.sha256(p.tobytes()).hexdigest(),
        }
    (HERE / "golden_c1.json").write_text(json.dumps(c1, indent=1) + "\n")
    print("wrote", [f.name for f in HERE.iterdir()])


if __name__ == "__main__":
    main()
