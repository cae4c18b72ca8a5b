"""ctypes access to the checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

Two CPU oracles are exposed:

* ``port``  -- oracle/_build/liboracle.so, the plain-C restatement of the reference's
  algorithm (oracle/radix_oracle.c).  Builds anywhere gcc exists.
* ``ref``   -- oracle/_ref/libref_sort.so, the UNMODIFIED reference header
  (/root/reference/radixSort.hpp) compiled by oracle/Makefile.  Needs an AVX-512
  (F/BW/DQ/VL/VBMI/VBMI2) host; ``ref_available()`` says whether it can run here.

A third, independent statement of the key order (map the key to an
unsigned integer whose ascending order is the reference's order) is given in numpy by
``order_key`` / ``total_order_sorted_keys``.

Nothing in the product imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
PORT_SO = ORACLE_DIR / "_build" / "liboracle.so"
REF_SO = ORACLE_DIR / "_ref" / "libref_sort.so"

KEY_DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64,
              np.float32, np.float64]
KEY_CODE = {np.dtype(d): i for i, d in enumerate(KEY_DTYPES)}


def key_code(dtype) -> int:
    return KEY_CODE[np.dtype(dtype)]


def build_port() -> Path:
    if not PORT_SO.exists() or PORT_SO.stat().st_mtime < (ORACLE_DIR / "radix_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)
    return PORT_SO


_port = None
_ref = None
_ref_checked = False


def port():
    global _port
    if _port is None:
        lib = ctypes.CDLL(str(build_port()))
        vp, i64, i32, u32p = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)
        vpp = ctypes.POINTER(ctypes.c_void_p)
        lib.oracle_sort_soa.argtypes = [vp, i32, i64, i32, i32, vpp, u32p, i64, i32]
        lib.oracle_sort_soa.restype = i32
        lib.oracle_sort_aos.argtypes = [vp, i32, ctypes.c_uint32, i64, i32, i64, i32]
        lib.oracle_sort_aos.restype = i32
        lib.oracle_gen_uniform_u32.argtypes = [vp, i64, ctypes.c_uint32]
        lib.oracle_gen_uniform_u32.restype = None
        lib.oracle_make_payloads.argtypes = [vp, i32, i64, i32, vpp, u32p]
        lib.oracle_make_payloads.restype = None
        lib.oracle_check_payloads.argtypes = [vp, i32, i64, i32, vpp, u32p]
        lib.oracle_check_payloads.restype = i32
        lib.oracle_is_sorted.argtypes = [vp, i32, ctypes.c_uint32, i64, i32]
        lib.oracle_is_sorted.restype = i32
        _port = lib
    return _port


def _cpu_has_ref_isa() -> bool:
    try:
        flags = Path("/proc/cpuinfo").read_text().split("flags", 1)[1].split("\n", 1)[0].split()
    except Exception:
        return False
    need = {"avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512vbmi", "avx512_vbmi2"}
    return need.issubset(flags)


def ref():
    """The compiled reference, or None when it is not built or this CPU cannot run it."""
    global _ref, _ref_checked
    if not _ref_checked:
        _ref_checked = True
        if REF_SO.exists() and _cpu_has_ref_isa():
            lib = ctypes.CDLL(str(REF_SO))
            vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
            lib.ref_cpu_ok.restype = i32
            if lib.ref_cpu_ok():
                lib.ref_sort_soa.argtypes = [vp, i32, i64, i32, i32, ctypes.POINTER(ctypes.c_void_p),
                                             ctypes.POINTER(ctypes.c_uint32)]
                lib.ref_sort_soa.restype = i32
                lib.ref_sort_aos.argtypes = [vp, i32, ctypes.c_uint32, i64, i32]
                lib.ref_sort_aos.restype = i32
                _ref = lib
    return _ref


def ref_available() -> bool:
    return ref() is not None


def _stream_args(payloads):
    n = len(payloads)
    ptrs = (ctypes.c_void_p * max(n, 1))(*[p.ctypes.data for p in payloads])
    sizes = (ctypes.c_uint32 * max(n, 1))(*[p.dtype.itemsize for p in payloads])
    return n, ptrs, sizes


def _check_arrays(keys, payloads):
    assert keys.flags.c_contiguous and keys.ndim == 1
    for p in payloads:
        assert p.flags.c_contiguous and p.ndim == 1 and len(p) == len(keys)


def port_sort_soa(keys: np.ndarray, payloads=(), up=True, thresh=16, cmp_sorter=0):
    """In place, like simd_sort::radix_sort::sort<Up>(thresh, num, keys, payloads...)."""
    _check_arrays(keys, payloads)
    n, ptrs, sizes = _stream_args(payloads)
    rc = port().oracle_sort_soa(keys.ctypes.data, key_code(keys.dtype), len(keys), int(up), n, ptrs, sizes,
                                thresh, cmp_sorter)
    assert rc == 0, rc


def port_sort_aos(records: np.ndarray, key_dtype, up=True, thresh=16, cmp_sorter=0):
    """records: (num, record_bytes) uint8, C-contiguous; key of key_dtype at byte offset 0."""
    assert records.dtype == np.uint8 and records.ndim == 2 and records.flags.c_contiguous
    rc = port().oracle_sort_aos(records.ctypes.data, key_code(key_dtype), records.shape[1], records.shape[0],
                                int(up), thresh, cmp_sorter)
    assert rc == 0, rc


def ref_sort_soa(keys: np.ndarray, payloads=(), up=True):
    _check_arrays(keys, payloads)
    n, ptrs, sizes = _stream_args(payloads)
    rc = ref().ref_sort_soa(keys.ctypes.data, key_code(keys.dtype), len(keys), int(up), n, ptrs, sizes)
    assert rc == 0, f"ref_sort_soa rc={rc} (shape not instantiated in oracle/ref_shim.cpp?)"


def ref_sort_aos(records: np.ndarray, key_dtype, up=True):
    assert records.dtype == np.uint8 and records.ndim == 2 and records.flags.c_contiguous
    rc = ref().ref_sort_aos(records.ctypes.data, key_code(key_dtype), records.shape[1], records.shape[0], int(up))
    assert rc == 0, rc


# ---------------------------------------------------------------------------------------------
# numpy statement of the key order (follows src/radix_sort.hpp:51-64)
# ---------------------------------------------------------------------------------------------
def order_key(keys: np.ndarray, up=True) -> np.ndarray:
    """Unsigned integers whose ascending order is the reference's order of `keys`."""
    dt = keys.dtype
    bits = dt.itemsize * 8
    udt = np.dtype(f"uint{bits}")
    u = keys.view(udt).copy()
    sign = udt.type(1 << (bits - 1))
    if dt.kind == "i":
        u ^= sign
    elif dt.kind == "f":
        neg = (u & sign) != 0
        u = np.where(neg, ~u, u ^ sign)
    if not up:
        u = ~u
    return u


def total_order_sorted_keys(keys: np.ndarray, up=True) -> np.ndarray:
    return keys[np.argsort(order_key(keys, up), kind="stable")]


def runs_multiset_equal(keys_sorted: np.ndarray, payloads_a, payloads_b) -> bool:
    """Payload parity rule: same (key, payload...) multiset inside every equal-key run.

    keys_sorted is the (identical) sorted key sequence of both results; payloads_x are lists of
    arrays (any itemsize; compared bytewise)."""
    n = len(keys_sorted)
    if n == 0:
        return True

    def rows(pl):
        cols = [np.ascontiguousarray(p).view(np.uint8).reshape(n, -1) for p in pl]
        return np.concatenate(cols, axis=1) if cols else np.zeros((n, 0), np.uint8)

    ra, rb = rows(payloads_a), rows(payloads_b)
    if ra.shape != rb.shape:
        return False
    if ra.shape[1] == 0:
        return True
    kb = keys_sorted.view(np.uint8).reshape(n, -1)
    # run id = index of first element of the run
    change = np.ones(n, bool)
    change[1:] = np.any(kb[1:] != kb[:-1], axis=1)
    run = np.cumsum(change) - 1

    def canon(r):
        # sort rows inside each run: lexsort by (payload bytes..., run)
        order = np.lexsort(tuple(r[:, j] for j in range(r.shape[1] - 1, -1, -1)) + (run,))
        return r[order]

    return bool(np.array_equal(canon(ra), canon(rb)))


# ---------------------------------------------------------------------------------------------
# Input distributions (shape of src/data.hpp:105-170; numpy RNG, so NOT bit-identical to the
# reference's std:: distributions -- only `c1_input` below is bit-identical).
# ---------------------------------------------------------------------------------------------
DISTRIBUTIONS = ["Gaussian", "Uniform", "Zero", "ZeroOne", "Sorted", "ReverseSorted", "AlmostSorted",
                 "AlmostReverseSorted"]


def _uniform(rng, dtype, n):
    dt = np.dtype(dtype)
    if dt.kind in "iu":
        return rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, size=n, dtype=dt, endpoint=True)
    return rng.uniform(-1.0, 1.0, size=n).astype(dt)


def _gaussian(rng, dtype, n):
    dt = np.dtype(dtype)
    if dt.kind in "iu":
        # the reference assigns round(N(0,100)) to K; wrap like a C conversion of an in-range value
        return np.round(rng.normal(0, 100, size=n)).astype(np.int64).astype(dt)
    return rng.normal(0, 1.0, size=n).astype(dt)


def make_keys(dist: str, dtype, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if dist == "Zero":
        return np.zeros(n, dt)
    if dist == "ZeroOne":
        return rng.integers(0, 2, size=n).astype(dt)
    if dist == "Uniform":
        return _uniform(rng, dt, n)
    if dist == "Gaussian":
        return _gaussian(rng, dt, n)
    base = _uniform(rng, dt, n) if dt.kind in "iu" else _gaussian(rng, dt, n)
    base = np.sort(base)
    if "Reverse" in dist:
        base = base[::-1].copy()
    if dist.startswith("Almost") and n > 0:
        for _ in range(int(2 ** np.log10(n))):
            i, j = rng.integers(0, n, size=2)
            base[i], base[j] = base[j], base[i]
    return base


def c1_input(n: int, seed: int = 42):
    """BASELINE.json config 1 input, bit-identical to Data<uint32_t,uint32_t>(n, Uniform, seed)."""
    keys = np.empty(n, np.uint32)
    pay = np.empty(n, np.uint32)
    port().oracle_gen_uniform_u32(keys.ctypes.data, n, seed)
    _, ptrs, sizes = _stream_args([pay])
    port().oracle_make_payloads(keys.ctypes.data, 4, n, 1, ptrs, sizes)
    return keys, pay


def check_payloads(keys: np.ndarray, payloads) -> bool:
    n, ptrs, sizes = _stream_args(list(payloads))
    return bool(port().oracle_check_payloads(keys.ctypes.data, keys.dtype.itemsize, len(keys), n, ptrs, sizes))


def is_sorted(keys: np.ndarray, up=True, stride=None, key_dtype=None) -> bool:
    kd = np.dtype(key_dtype or keys.dtype)
    stride = stride or kd.itemsize
    num = keys.nbytes // stride
    return bool(port().oracle_is_sorted(keys.ctypes.data, key_code(kd), stride, num, int(up)))


def make_records(keys: np.ndarray, record_bytes: int) -> np.ndarray:
    """AoS records (num, record_bytes) uint8: key at offset 0, then key-derived payload bytes
    (one opaque payload of record_bytes - sizeof(key) bytes, src/data.hpp:393-406)."""
    n, kb = len(keys), keys.dtype.itemsize
    rec = np.zeros((n, record_bytes), np.uint8)
    rec[:, :kb] = np.ascontiguousarray(keys).view(np.uint8).reshape(n, kb)
    if record_bytes > kb and n > 0:
        flat = np.zeros(n * (record_bytes - kb), np.uint8)
        ptrs = (ctypes.c_void_p * 1)(flat.ctypes.data)
        sizes = (ctypes.c_uint32 * 1)(record_bytes - kb)
        keys_c = np.ascontiguousarray(keys)
        port().oracle_make_payloads(keys_c.ctypes.data, kb, n, 1, ptrs, sizes)
        rec[:, kb:] = flat.reshape(n, -1)
    return rec


# ---------------------------------------------------------------------------------------------------
# counter-based key generator shared by host (numpy) and device (torch) code: key i = mix64(seed + i)
# (splitmix64 finaliser; the same function sortbench.cu and bench.py use), so that both arms of a
# comparison sort the very same records without shipping them.
# ---------------------------------------------------------------------------------------------------
_M1, _M2, _M3 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def mix64_numpy(start: int, count: int, seed: int) -> np.ndarray:
    """uint64 keys mix64(seed + i) for i in [start, start + count)"""
    with np.errstate(over="ignore"):
        x = np.arange(start, start + count, dtype=np.uint64) + np.uint64(seed % (1 << 64))
        x = x + np.uint64(_M1)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(_M2)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(_M3)
        return x ^ (x >> np.uint64(31))


def mix64_torch(start: int, count: int, seed: int, device):
    """the same keys on a torch device, as an int64 tensor holding the uint64 bit patterns"""
    import torch

    def s64(v):  # python int -> the int64 with the same low 64 bits
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(t, s):
        return (t >> s) & ((1 << (64 - s)) - 1)

    x = torch.arange(start, start + count, dtype=torch.int64, device=device) + s64(seed)
    x = x + s64(_M1)
    x = (x ^ lsr(x, 30)) * s64(_M2)
    x = (x ^ lsr(x, 27)) * s64(_M3)
    return x ^ lsr(x, 31)
