"""GPU parity tests at DEFAULT options and production-flow sizes (-m gpu).

The tests of test_gpu_parity.py reach the large-sort flow (plan read-back, sparse probe sampling, unstable first
pass, tile-local ordering in the last pass, junction kernel) by forcing it at small sizes; these run it the
way a user does: no options touched, n >= 2^25, the BASELINE.json config shapes.  Oracles: numpy's total
order (tests/oracle_lib.py: order_key), per-record permutation checks, and for config 2 the SHA-256 of
the compiled reference's sorted key stream (tests/golden/golden_c2.json, tests/golden/make_golden_c2.py).
"""
import hashlib
import json

import numpy as np
import pytest

import oracle_lib as O
import simd_radix_sort_b200 as S

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N25 = 1 << 25


def dev(a):
    return torch.from_numpy(a).cuda()


def test_c2_shape_default_options_2p25(golden_dir):
    """u64 key + u64 payload, uniform (BASELINE.json configs[1] shape), generator shared with the reference run"""
    n = N25
    gold = json.loads((golden_dir / "golden_c2.json").read_text())[str(n)]
    keys0 = O.mix64_torch(0, n, gold["seed"], "cuda")
    k = keys0.clone()
    p = torch.arange(n, dtype=torch.int64, device="cuda")
    S.sort(n, k.view(torch.uint64), p.view(torch.uint64), up=True)
    torch.cuda.synchronize()
    st = S.last_stats()
    assert st["algo"] == 2 and st["fell_back"] == 0 and st["segfix_passes"] == 0, st  # the production flow ran
    hk = k.cpu().numpy().view(np.uint64)
    assert hashlib.sha256(hk.tobytes()).hexdigest() == gold["sha256_sorted_keys"]  # == the compiled reference
    assert hk[:3].tolist() == gold["head"] and int(hk[-1]) == gold["tail"]
    assert bool((keys0[p] == k).all().item())                      # every payload followed its key
    assert bool((torch.sort(p).values == torch.arange(n, device="cuda")).all().item())  # ... and is a permutation


def test_u32_u32_default_options_2p25_wide_tiles():
    """config 1's shape (uint32 key + uint32 payload) at 2^25: the automatically chosen 8192-key tiles, plan read-back,
    unstable first pass"""
    n = N25
    rng = np.random.default_rng(11)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    k, p = dev(keys), dev(np.arange(n, dtype=np.uint32))
    S.sort(n, k, p, up=True)
    torch.cuda.synchronize()
    hk, hp = k.cpu().numpy(), p.cpu().numpy()
    assert hk.tobytes() == np.sort(keys).tobytes()
    assert keys[hp].tobytes() == hk.tobytes() and np.array_equal(np.sort(hp), np.arange(n, dtype=np.uint32))


def test_c3_shape_default_options_2p25():
    """f32 key + (i32, f64, u16) payload streams, descending, with the SURVEY 8d edge set (configs[2] shape)"""
    n = N25
    rng = np.random.default_rng(3)
    keys = rng.uniform(-1, 1, size=n).astype(np.float32)
    keys[:8 * 1_000_003:1_000_003] = [0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, -3.4e38]
    i32 = np.arange(n, dtype=np.int32)
    f64 = np.arange(n, dtype=np.float64)
    u16 = (np.arange(n) % 65536).astype(np.uint16)
    k, a, b, c = dev(keys), dev(i32), dev(f64), dev(u16)
    S.sort(n, k, a, b, c, up=False)
    torch.cuda.synchronize()
    hk, ha = k.cpu().numpy(), a.cpu().numpy()
    assert hk.tobytes() == O.total_order_sorted_keys(keys, False).tobytes()
    assert keys[ha].tobytes() == hk.tobytes()
    assert np.array_equal(np.sort(ha), i32)
    assert np.array_equal(b.cpu().numpy(), ha.astype(np.float64))
    assert np.array_equal(c.cpu().numpy(), (ha % 65536).astype(np.uint16))


@pytest.mark.parametrize("dist", ["zipf", "few_unique", "zero_one"])
def test_c4_shape_default_options_2p25(dist):
    """combined DataElement<int64, double> records (configs[3] shape): Zipf-skewed, few-unique (-8..7) and the
    reference's ZeroOne distribution (src/data.hpp:110-119)"""
    n = N25
    rng = np.random.default_rng(4)
    if dist == "zipf":
        ranks = np.minimum(rng.zipf(1.2, size=n), 1 << 20) - 1
        table = rng.integers(-2**63, 2**63 - 1, size=1 << 20, dtype=np.int64)
        keys = table[ranks]
    elif dist == "few_unique":
        keys = rng.integers(-8, 8, size=n, dtype=np.int64)
    else:
        keys = rng.integers(0, 2, size=n, dtype=np.int64)
    rec = np.empty((n, 2), np.int64)
    rec[:, 0] = keys
    rec[:, 1] = np.arange(n, dtype=np.float64).view(np.int64)
    r = dev(rec.view(np.uint8).reshape(n, 16))
    S.sort_combined(n, r, np.int64, up=True)
    torch.cuda.synchronize()
    out = r.cpu().numpy().view(np.int64).reshape(n, 2)
    ok = np.ascontiguousarray(out[:, 0])
    op = np.ascontiguousarray(out[:, 1]).view(np.float64).astype(np.int64)
    assert ok.tobytes() == np.sort(keys).tobytes()
    assert np.array_equal(keys[op], ok)                       # the payload travelled with its key
    assert np.array_equal(np.sort(op), np.arange(n))          # ... and every record is still there, once


def test_more_than_2p32_records_u16_key_u8_payload():
    """SortIndex is 64-bit (src/common.hpp:14): n > 2^32.  Checked by histogram (the sorted key sequence is
    determined by the key counts), sortedness and payload = f(key)."""
    n = (1 << 32) + 12_345
    chunk = 1 << 28
    keys = torch.empty(n, dtype=torch.int16, device="cuda")
    pay = torch.empty(n, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(65536, dtype=torch.int64, device="cuda")
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        x = O.mix64_torch(s, c, 99, "cuda")
        kc = ((x >> 20) & 0xFFFF)
        kc = torch.where((x & 7) == 0, kc & 0xFF, kc)          # some skew: an eighth of the keys share 256 values
        counts += torch.bincount(kc, minlength=65536)
        keys[s:s + c] = (kc - 32768).to(torch.int16)            # signed keys -32768..32767
        pay[s:s + c] = ((kc * 7 + 3) & 0xFF).to(torch.uint8)
        del x, kc
    S.sort(n, keys, pay, up=False)
    torch.cuda.synchronize()
    # descending: the key sequence must be 32767 x counts[65535], 32766 x counts[65534], ...
    ends = torch.cumsum(counts.flip(0), 0)
    got = torch.zeros(65536, dtype=torch.int64, device="cuda")
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        kc = keys[s:s + c].to(torch.int64) + 32768
        got += torch.bincount(kc, minlength=65536)
        assert bool((kc[1:] <= kc[:-1]).all().item())
        if s > 0:
            assert int(keys[s - 1].item()) >= int(keys[s].item())
        assert bool((pay[s:s + c] == ((kc * 7 + 3) & 0xFF).to(torch.uint8)).all().item())
        del kc
    assert bool((got == counts).all().item())
    assert int(ends[-1].item()) == n


def test_c2_full_size_1e9_matches_the_compiled_reference(golden_dir):
    """BASELINE.json configs[1] at its stated size: 1e9 u64+u64 records, default options, device-resident; the
    sorted key stream must hash to what the compiled reference produced for the same records."""
    n = 1_000_000_000
    gold = json.loads((golden_dir / "golden_c2.json").read_text()).get(str(n))
    if gold is None:
        pytest.skip("golden_c2.json has no 1e9 entry")
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * (1 << 30):
        pytest.skip("needs 60 GB of device memory")
    step = 1 << 27
    k = torch.empty(n, dtype=torch.int64, device="cuda")
    for s in range(0, n, step):
        c = min(step, n - s)
        k[s:s + c] = O.mix64_torch(s, c, gold["seed"], "cuda")
    p = torch.arange(n, dtype=torch.int64, device="cuda")
    S.sort(n, k.view(torch.uint64), p.view(torch.uint64), up=True)
    torch.cuda.synchronize()
    st = S.last_stats()
    assert st["algo"] == 2 and st["fell_back"] == 0, st
    h = hashlib.sha256()
    for s in range(0, n, step):
        c = min(step, n - s)
        h.update(k[s:s + c].cpu().numpy().tobytes())
        # payloads: regenerate the key every payload points at
        src = p[s:s + c]
        # mix64(seed + i) for arbitrary i: the generator is counter-based
        x = src + gold["seed"]
        def s64(v):
            v &= (1 << 64) - 1
            return v - (1 << 64) if v >= (1 << 63) else v
        def lsr(t, sh):
            return (t >> sh) & ((1 << (64 - sh)) - 1)
        x = x + s64(O._M1)
        x = (x ^ lsr(x, 30)) * s64(O._M2)
        x = (x ^ lsr(x, 27)) * s64(O._M3)
        x = x ^ lsr(x, 31)
        assert bool((x == k[s:s + c]).all().item())
    assert h.hexdigest() == gold["sha256_sorted_keys"]
    # every record still there exactly once: the payloads are a permutation of 0..n-1 (sum and sum of squares mod 2^64)
    assert int(p.sum().item()) == n * (n - 1) // 2
