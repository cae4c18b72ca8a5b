#!/usr/bin/env python
"""perf_dat.py -- SURVEY.md 8(f) rank 4: the reference's performance-harness OUTPUT, regenerated with a B200
column (TEST / MEASUREMENT INFRASTRUCTURE: it may time the compiled reference under oracle/_ref, so it lives
under tests/, not in the product).

Mirrors /root/reference/src/perf.hpp:
  * `tpe-<key>[-<payloads>]-<Distribution>.dat`    (perfTestNum, perf.hpp:370-414): header
        number_of_elements <method> <method> ...
    then one row per num = 1, 2, 4, ... 2^22 (here up to --max-log2) with nanoseconds per element, %.6f;
  * `<key>[-<payloads>]-<Distribution>-262144.dat` (perfTest, perf.hpp:416-461): header
        sort_method nanoseconds_per_element
    one row per method, an empty line at the end.
Method (perf.hpp:66-88): max(1, 2^18/num) warm-up sorts, then the mean over max(1, 2^22/num) sorts, every one
on a fresh copy of the input; the distributions are those of src/data.hpp:105-170 (tests/oracle_lib.make_keys).
Columns:
  RadixB200      arrays resident in device memory, CUDA-event time of the sort call (the perf path);
  RadixB200Host  the drop-in call on HOST arrays (H2D + sort + D2H inside, wall clock) -- what a caller of the
                 reference's sort<>() gets without touching their code;
  RadixSIMD      the unmodified reference (oracle/_ref, one thread, CLOCK_PROCESS_CPUTIME-like wall clock of
                 the call) on the same host -- only when it can run here.

    python tests/perf_dat.py --out /tmp/radixSortData-b200 [--max-log2 22] [--types int32-int32,int64-int64,...]
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
import oracle_lib as O  # noqa: E402

TYPE_NAMES = {"uint8": np.uint8, "int8": np.int8, "uint16": np.uint16, "int16": np.int16, "uint32": np.uint32,
              "int32": np.int32, "uint64": np.uint64, "int64": np.int64, "float": np.float32, "double": np.float64}


def reps(num: int):
    return max(1, (1 << 18) // num), max(1, (1 << 22) // num)   # warm-ups, timed tests (perf.hpp:69-70)


def make_input(desc: str, dist: str, num: int, seed: int):
    names = desc.split("-")
    keys = O.make_keys(dist, TYPE_NAMES[names[0]], num, seed)
    pays = [(np.arange(num) % 251).astype(TYPE_NAMES[p]) for p in names[1:]]
    return keys, pays


def time_b200_device(S, torch, keys, pays, n_warm, n_test):
    dk0 = torch.from_numpy(keys).cuda()
    dp0 = [torch.from_numpy(p).cuda() for p in pays]
    dk, dp = torch.empty_like(dk0), [torch.empty_like(p) for p in dp0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for i in range(n_warm + n_test):
        dk.copy_(dk0)
        for a, b in zip(dp, dp0):
            a.copy_(b)
        e0.record()
        S.sort(len(keys), dk, *dp, up=True)
        e1.record()
        e1.synchronize()
        if i >= n_warm:
            total += e0.elapsed_time(e1) * 1e6
    got = dk.cpu().numpy()
    assert got.tobytes() == O.total_order_sorted_keys(keys, True).tobytes(), "RadixB200: not sorted"
    return total / n_test / max(len(keys), 1)


def time_host(fn, keys, pays, n_warm, n_test, check):
    total = 0.0
    k = p = None
    for i in range(n_warm + n_test):
        k, p = keys.copy(), [x.copy() for x in pays]
        t0 = time.perf_counter_ns()
        fn(k, p)
        dt = time.perf_counter_ns() - t0
        if i >= n_warm:
            total += dt
    if check:
        assert k.tobytes() == O.total_order_sorted_keys(keys, True).tobytes(), "not sorted"
    return total / n_test / max(len(keys), 1)


def run(out: Path, descs, dists, max_log2: int, seed: int, max_reps: int):
    import torch
    import simd_radix_sort_b200 as S

    out.mkdir(parents=True, exist_ok=True)
    have_ref = O.ref_available()
    methods = ["RadixB200", "RadixB200Host"] + (["RadixSIMD"] if have_ref else [])

    def measure(desc, dist, num):
        keys, pays = make_input(desc, dist, num, seed)
        n_warm, n_test = reps(num)
        n_warm, n_test = min(n_warm, max_reps), min(n_test, max_reps)
        row = [time_b200_device(S, torch, keys, pays, n_warm, n_test),
               time_host(lambda k, p: S.sort(len(k), k, *p, up=True), keys, pays, min(n_warm, 8), min(n_test, 32), True)]
        if have_ref:
            row.append(time_host(lambda k, p: O.ref_sort_soa(k, p, True), keys, pays, min(n_warm, 8), min(n_test, 32), False))
        return row

    written = []
    for desc in descs:
        for dist in dists:
            f = out / f"tpe-{desc}-{dist}.dat"
            with f.open("w") as fh:
                fh.write("number_of_elements " + " ".join(methods) + "\n")
                for lg in range(0, max_log2 + 1):
                    num = 1 << lg
                    fh.write(str(num) + "".join(f" {v:.6f}" for v in measure(desc, dist, num)) + "\n")
            written.append(f)
            num = 1 << 18
            f = out / f"{desc}-{dist}-{num}.dat"
            with f.open("w") as fh:
                fh.write("sort_method nanoseconds_per_element\n")
                for name, v in zip(methods, measure(desc, dist, num)):
                    fh.write(f"{name} {v:.6f}\n")
                fh.write("\n")
            written.append(f)
    return written


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/radixSortData-b200")
    ap.add_argument("--types", default="int32-int32,int64-int64,float-int32,double-int64,int32")
    ap.add_argument("--dists", default="Uniform,Gaussian,Zero,ZeroOne,Sorted,AlmostSorted")
    ap.add_argument("--max-log2", type=int, default=22)
    ap.add_argument("--max-reps", type=int, default=64, help="cap on warm-ups / timed sorts per size (the reference uses up to 2^22)")
    ap.add_argument("--seed", type=int, default=42)
    a = ap.parse_args()
    files = run(Path(a.out), a.types.split(","), a.dists.split(","), a.max_log2, a.seed, a.max_reps)
    print("\n".join(str(f) for f in files))


if __name__ == "__main__":
    main()
