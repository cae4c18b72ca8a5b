"""CPU tests that pin the oracle (oracle/radix_oracle.c) -- see its header.

Pins, in order of authority:
 1. the compiled reference itself (oracle/_ref, only where /root/reference was compiled and the
    CPU has AVX-512 VBMI2 -- i.e. in the build container),
 2. tests/golden/*.npz + golden_c1.json, produced by that compiled reference (travel everywhere),
 3. the known-answer values of SURVEY.md section 4.
"""
import hashlib
import json

import numpy as np
import pytest

import oracle_lib as O

ALL_DTYPES = [np.dtype(d) for d in O.KEY_DTYPES]


def test_kat_c1_generator_and_sort(golden_dir):
    gold = json.loads((golden_dir / "golden_c1.json").read_text())
    for n_str, g in gold.items():
        n = int(n_str)
        k, p = O.c1_input(n, 42)
        assert k[:4].tolist() == [1608637542, 3421126067, 4083286876, 787846414]
        assert p[:2].tolist() == [1007292957, 498833468]
        assert k[:4].tolist() == g["input_first_keys"] and p[:2].tolist() == g["input_first_payloads"]
        O.port_sort_soa(k, [p], True)
        assert k[:3].tolist() == g["sorted_head"] and int(k[-1]) == g["sorted_tail"]
        assert hashlib.sha256(k.tobytes()).hexdigest() == g["sha256_sorted_keys"]
        # payload is a function of the key (src/data.hpp:393-406) so the payload stream is pinned too
        assert hashlib.sha256(p.tobytes()).hexdigest() == g["sha256_sorted_payloads"]
        assert O.check_payloads(k, [p]) and O.is_sorted(k, True)


def test_port_matches_golden_soa(golden_dir):
    z = np.load(golden_dir / "soa_keys.npz")
    n_cases = 0
    for name in z.files:
        if not name.startswith("in|"):
            continue
        _, dt, up, dist, n = name.split("|")
        keys = z[name].copy()
        want = z["out|" + name[3:]]
        idx = np.arange(len(keys), dtype=np.uint32)
        O.port_sort_soa(keys, [idx], bool(int(up)))
        assert keys.tobytes() == want.tobytes(), name
        assert np.array_equal(np.sort(idx), np.arange(len(keys))), name  # a permutation
        assert z[name][idx].tobytes() == keys.tobytes(), name  # payload moved with its key
        if int(n) > 16 or dist != "Specials":
            assert O.total_order_sorted_keys(z[name], bool(int(up))).tobytes() == want.tobytes(), name
        n_cases += 1
    assert n_cases >= 10 * 2 * 6 * 7


def test_port_matches_golden_aos(golden_dir):
    z = np.load(golden_dir / "aos_records.npz")
    n_cases = 0
    for name in z.files:
        if not name.startswith("inkeys|"):
            continue
        _, dt, rb, up, dist, n = name.split("|")
        keys = z[name]
        rec = O.make_records(keys, int(rb))
        O.port_sort_aos(rec, np.dtype(dt), bool(int(up)))
        want = O.make_records(z["outkeys|" + name[7:]], int(rb))
        assert rec.tobytes() == want.tobytes(), name
        n_cases += 1
    assert n_cases > 100


@pytest.mark.skipif(not O.ref_available(), reason="compiled reference (oracle/_ref) not runnable here")
@pytest.mark.parametrize("dt", ALL_DTYPES, ids=lambda d: d.name)
def test_port_matches_compiled_reference(dt):
    shapes = [[], [np.uint8], [np.uint32], [np.uint64, np.uint8], [np.uint32, np.uint64, np.uint16]]
    for up in (True, False):
        for dist in O.DISTRIBUTIONS:
            for n in (1, 10, 100, 3000):
                keys = O.make_keys(dist, dt, n, seed=n + 5)
                for shape in shapes:
                    pay = [np.arange(n).astype(s) for s in shape]
                    k1, p1 = keys.copy(), [p.copy() for p in pay]
                    k2, p2 = keys.copy(), [p.copy() for p in pay]
                    O.ref_sort_soa(k1, p1, up)
                    O.port_sort_soa(k2, p2, up)
                    assert k1.tobytes() == k2.tobytes(), (dist, n, up, shape)
                    assert O.runs_multiset_equal(k1, p1, p2), (dist, n, up, shape)


@pytest.mark.skipif(not O.ref_available(), reason="compiled reference (oracle/_ref) not runnable here")
def test_port_matches_compiled_reference_aos():
    for dt in ALL_DTYPES:
        rb = dt.itemsize
        while rb <= 64:
            for up in (True, False):
                keys = O.make_keys("Gaussian", dt, 500, seed=rb)
                r1 = O.make_records(keys, rb)
                r2 = r1.copy()
                O.ref_sort_aos(r1, dt, up)
                O.port_sort_aos(r2, dt, up)
                assert r1.tobytes() == r2.tobytes()
            rb *= 2


def test_order_key_is_the_reference_order():
    # float specials incl. -0.0 < +0.0 and +-inf (n > 16 so the sign bit is split first)
    for dt in (np.float32, np.float64):
        a = np.array([0.0, -0.0, np.inf, -np.inf, 1.0, -1.0, 5e-324, -5e-324] * 4, dtype=dt)
        for up in (True, False):
            k = a.copy()
            O.port_sort_soa(k, [], up)
            assert k.tobytes() == O.total_order_sorted_keys(a, up).tobytes()


def test_small_and_empty():
    for n in (0, 1):
        k = np.arange(n, dtype=np.int32)
        O.port_sort_soa(k, [], True)
        assert len(k) == n


def test_nosort_threshold_semantics():
    # CmpSorterNoSort (src/cmp_sorters.hpp:66-78): buckets of <= thresh elements stay unsorted, but every
    # element is within its final bucket: sorting each aligned radix bucket completes the sort.
    keys = O.make_keys("Uniform", np.uint16, 5000, seed=3)
    k = keys.copy()
    O.port_sort_soa(k, [], True, thresh=16, cmp_sorter=1)
    assert sorted(k.tolist()) == sorted(keys.tolist())
    full = np.sort(keys)
    # an element may be displaced by fewer than thresh positions
    assert np.all(np.abs(np.searchsorted(full, k, side="left") - np.arange(len(k))) < 16 + 16)
