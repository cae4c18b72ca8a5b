"""World-size-2 `gloo` test (CPU) of the host-side logic of the multi-GPU shard path:
top-bits histogram -> all-reduce -> splitters (C ABI: b200sort_mgpu_splitters) -> exchange plan
(b200sort_mgpu_plan) -> all-to-all.  The record movement itself is emulated with numpy here (the CUDA
partition/exchange kernels need GPUs: see tests/test_gpu_mgpu.py); what is checked is that every rank
derives the same splitters, that the plan's counts match the data, and that after the exchange the
ranks hold disjoint, ordered key ranges whose union is the input."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent('''
    import ctypes, os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["B200_ROOT"]); sys.path.insert(0, os.path.join(os.environ["B200_ROOT"], "tests"))
    import simd_radix_sort_b200 as S
    import oracle_lib as O

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["B200_PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    L = S.lib()
    u64p, u32p = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)
    for case, (dt, up) in enumerate([(np.uint64, True), (np.float32, False), (np.int16, True)]):
        rng = np.random.default_rng(100 * case + rank)
        n_local = 50_000 + 1000 * rank
        keys = O.make_keys("Uniform" if case != 1 else "Gaussian", dt, n_local, seed=7 * case + rank)
        bits = min(16, 8 * np.dtype(dt).itemsize)
        top = (O.order_key(keys, up) >> np.array(8 * np.dtype(dt).itemsize - bits).astype(O.order_key(keys, up).dtype)).astype(np.int64)
        local_hist = np.bincount(top, minlength=1 << bits).astype(np.uint64)
        g = torch.from_numpy(local_hist.astype(np.int64))
        dist.all_reduce(g)
        global_hist = g.numpy().astype(np.uint64)
        bounds = np.zeros(world + 1, np.uint32)
        assert L.b200sort_mgpu_splitters(global_hist.ctypes.data_as(u64p), bits, world, bounds.ctypes.data_as(u32p)) == 0
        # identical splitters everywhere
        allb = [None] * world
        dist.all_gather_object(allb, bounds.tolist())
        assert all(b == allb[0] for b in allb), allb
        send = np.zeros(world, np.uint64)
        assert L.b200sort_mgpu_plan(local_hist.ctypes.data_as(u64p), bits, world, bounds.ctypes.data_as(u32p),
                                    send.ctypes.data_as(u64p)) == 0
        dest = np.searchsorted(bounds[1:], top, side="right")
        assert np.array_equal(np.bincount(dest, minlength=world).astype(np.uint64), send)
        # exchange (emulated all-to-all-v)
        parts = [keys[dest == r] for r in range(world)]
        gathered = [None] * world
        dist.all_gather_object(gathered, parts)
        mine = np.concatenate([gathered[src][rank] for src in range(world)])
        mine = O.total_order_sorted_keys(mine, up)
        edges = [None] * world
        dist.all_gather_object(edges, (len(mine), O.order_key(mine[:1], up).tolist(), O.order_key(mine[-1:], up).tolist()))
        total = sum(e[0] for e in edges)
        all_n = [None] * world
        dist.all_gather_object(all_n, n_local)
        assert total == sum(all_n)
        for r in range(world - 1):
            if edges[r][0] and edges[r + 1][0]:
                assert edges[r][2][0] <= edges[r + 1][1][0], "rank ranges overlap"
        # loads are balanced to bin granularity plus the 1/64 share a boundary may give up for an aligned bin index
        assert max(e[0] for e in edges) <= total / world * (1 + 1 / 32) + global_hist.max() + 1
    dist.barrier()
    dist.destroy_process_group()
    print("MGPU_CPU_OK", rank)
''')


def test_gloo_world2_splitters_plan_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", B200_ROOT=str(ROOT), B200_PORT=port,
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"MGPU_CPU_OK {rank}" in out, out[-3000:]


# ------------------------------------------------------------------------------------------------
# Skewed keys: heavy histogram bins are refined 16 bits per level down to single key values, and
# the keys equal to such a value are divided by source rank and position block.  World-size-2 gloo test of the
# host logic through the C ABI (b200sort_mgpu_refine_splitters with a numpy histogram callback that all-reduces
# over gloo, b200sort_mgpu_tie_thresholds), on the distributions of src/data.hpp:105-170 that break bin-granular
# splitters: Zero, ZeroOne, few-unique(16), Zipf, and Gaussian doubles.
# ------------------------------------------------------------------------------------------------
WORKER_SKEW = textwrap.dedent('''
    import ctypes, os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["B200_ROOT"]); sys.path.insert(0, os.path.join(os.environ["B200_ROOT"], "tests"))
    import simd_radix_sort_b200 as S
    from simd_radix_sort_b200 import _api
    import oracle_lib as O

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["B200_PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    L = S.lib()
    NB = 256          # position blocks of the tie split (the library uses 1024)
    VW = 8            # virtual world: every gloo rank plays VW // world ranks of an 8-GPU box
    per = VW // world
    n_local = 40_000
    rng = np.random.default_rng(5)
    table = rng.integers(-2**63, 2**63 - 1, size=1 << 12, dtype=np.int64)
    cases = {
        "zero": lambda r: np.zeros(n_local, np.int64),
        "zero_one": lambda r: np.random.default_rng(r).integers(0, 2, size=n_local).astype(np.uint64),
        "few_unique16": lambda r: np.random.default_rng(r).integers(-8, 8, size=n_local, dtype=np.int64),
        "zipf": lambda r: table[np.minimum(np.random.default_rng(r).zipf(1.3, size=n_local), 1 << 12) - 1],
        "gauss_f64": lambda r: np.random.default_rng(r).normal(0, 1, size=n_local),
        "uniform_u32": lambda r: np.random.default_rng(r).integers(0, 2**32, size=n_local, dtype=np.uint32),
        "one_heavy_value": lambda r: np.where(np.random.default_rng(r).random(n_local) < 0.45, np.uint64(77),
                                              np.random.default_rng(r + 50).integers(0, 2**64, size=n_local, dtype=np.uint64)),
    }
    for name, gen in cases.items():
        for up in (True, False):
            mine = [np.ascontiguousarray(gen(rank * per + q)) for q in range(per)]     # my virtual ranks' keys
            kb = mine[0].dtype.itemsize
            okeys = [O.order_key(k, up).astype(np.uint64) for k in mine]               # ordered-key space

            def hist_cb(ctx, n_ranges, lo, shift, nb, out):
                h = np.zeros((n_ranges, 65536), np.int64)
                for j in range(n_ranges):
                    for u in okeys:
                        sel = u >= np.uint64(lo[j])
                        b = ((u[sel] - np.uint64(lo[j])) >> np.uint64(shift[j]))
                        b = b[b < np.uint64(nb[j])].astype(np.int64)
                        h[j] += np.bincount(b, minlength=65536)
                t = torch.from_numpy(h)
                dist.all_reduce(t)
                res = np.ascontiguousarray(t.numpy().astype(np.uint64))   # (kept alive across the memmove)
                ctypes.memmove(out, res.ctypes.data, n_ranges * 65536 * 8)
                return 0

            total = VW * n_local
            keys_out = np.zeros(VW - 1, np.uint64)
            tie_out = np.zeros(VW - 1, np.uint32)
            cb = _api.HIST_FN(hist_cb)
            rc = L.b200sort_mgpu_refine_splitters(VW, kb, total, cb, None, keys_out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                  tie_out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
            assert rc == 0, L.b200sort_last_error()
            assert np.all(keys_out[1:] >= keys_out[:-1]), (name, keys_out)
            # position-block thresholds of every virtual rank for the tie splitters
            blk_shift = int(np.ceil(np.log2(max(n_local / NB, 1))))
            blk = {v: np.zeros(VW - 1, np.uint32) for v in range(rank * per, rank * per + per)}
            for tv in np.unique(keys_out[tie_out != 0]):
                less_l = np.array([int(np.sum(u < tv)) for u in okeys], np.int64)
                eq_l = np.zeros((per, NB), np.int64)
                for q, u in enumerate(okeys):
                    pos = np.nonzero(u == tv)[0] >> blk_shift
                    eq_l[q] = np.bincount(pos, minlength=NB)[:NB]
                gl = [None] * world
                dist.all_gather_object(gl, (less_l, eq_l))
                less_total = int(sum(g[0].sum() for g in gl))
                eq_all = np.concatenate([g[1] for g in gl]).astype(np.uint32)           # [VW][NB]
                which = np.nonzero((keys_out == tv) & (tie_out != 0))[0]
                targets = np.array([(total * (int(r) + 1)) // VW for r in which], np.uint64)
                for v in blk:
                    out = np.zeros(len(which), np.uint32)
                    rc = L.b200sort_mgpu_tie_thresholds(VW, v, NB, less_total, eq_all.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                                        len(which), targets.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                        out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
                    assert rc == 0
                    blk[v][which] = out
            # destination of every record: upper bound over the (key, block) pairs, like part_dest() on the device
            sizes = np.zeros(VW, np.int64)
            parts = []
            for q, u in enumerate(okeys):
                v = rank * per + q
                b = (np.arange(len(u)) >> blk_shift).astype(np.uint32)
                dest = np.zeros(len(u), np.int64)
                for r in range(VW - 1):
                    dest += ((u > keys_out[r]) | ((u == keys_out[r]) & (b >= blk[v][r]))).astype(np.int64)
                sizes += np.bincount(dest, minlength=VW)
                parts.append((u, dest))
            t = torch.from_numpy(sizes.copy())
            dist.all_reduce(t)
            sizes_all = t.numpy()
            assert sizes_all.sum() == total
            # no rank above capacity 1.125 * N / G  (the bar of VERDICT r1 item 4)
            assert sizes_all.max() <= 1.125 * total / VW, (name, up, sizes_all.tolist(), keys_out.tolist(), tie_out.tolist())
            # destinations are monotonic in the key: max key of rank d <= min key of rank d+1
            lo_hi = np.full((VW, 2), -1, np.float64)
            mins = np.full(VW, np.iinfo(np.uint64).max, np.uint64); maxs = np.zeros(VW, np.uint64); has = np.zeros(VW, bool)
            for u, dest in parts:
                for d in range(VW):
                    sel = dest == d
                    if sel.any():
                        has[d] = True
                        mins[d] = min(mins[d], u[sel].min()); maxs[d] = max(maxs[d], u[sel].max())
            g = [None] * world
            dist.all_gather_object(g, (mins, maxs, has))
            mins = np.min([x[0] for x in g], axis=0); maxs = np.max([x[1] for x in g], axis=0); has = np.any([x[2] for x in g], axis=0)
            last = None
            for d in range(VW):
                if has[d]:
                    if last is not None:
                        assert maxs[last] <= mins[d], (name, up, d)
                    last = d
    dist.barrier()
    dist.destroy_process_group()
    print("MGPU_SKEW_OK", rank)
''')


def test_gloo_world2_refined_splitters_and_tie_split(tmp_path):
    script = tmp_path / "worker_skew.py"
    script.write_text(WORKER_SKEW)
    port = str(31500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", B200_ROOT=str(ROOT), B200_PORT=port,
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"MGPU_SKEW_OK {rank}" in out, out[-3000:]


# ------------------------------------------------------------------------------------------------
# single-process unit tests of the same host logic: other world sizes (not a power of two), 1-/2-/4-byte keys
# ------------------------------------------------------------------------------------------------
def _split_and_count(keys_per_rank, world, n_blocks=64):
    """runs b200sort_mgpu_refine_splitters + b200sort_mgpu_tie_thresholds over in-memory 'ranks' and returns the
    number of records every destination gets (destination = upper bound over the (key, block) pairs)"""
    import ctypes
    import numpy as np
    import simd_radix_sort_b200 as S
    from simd_radix_sort_b200 import _api
    L = S.lib()
    kb = keys_per_rank[0].dtype.itemsize
    total = sum(len(k) for k in keys_per_rank)
    okeys = [k.astype(np.uint64) for k in keys_per_rank]      # (unsigned keys: their own ordered form)
    keep = []

    def hist_cb(ctx, n_ranges, lo, shift, nb, out):
        h = np.zeros((n_ranges, 65536), np.uint64)
        for j in range(n_ranges):
            for u in okeys:
                sel = u >= np.uint64(lo[j])
                b = (u[sel] - np.uint64(lo[j])) >> np.uint64(shift[j])
                b = b[b < np.uint64(nb[j])].astype(np.int64)
                h[j] += np.bincount(b, minlength=65536).astype(np.uint64)
        keep.append(h)
        ctypes.memmove(out, h.ctypes.data, n_ranges * 65536 * 8)
        return 0

    ns = world - 1
    keys_out, tie_out = np.zeros(max(ns, 1), np.uint64), np.zeros(max(ns, 1), np.uint32)
    rc = L.b200sort_mgpu_refine_splitters(world, kb, total, _api.HIST_FN(hist_cb), None,
                                          keys_out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                          tie_out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    assert rc == 0, L.b200sort_last_error()
    keys_out, tie_out = keys_out[:ns], tie_out[:ns]
    assert np.all(keys_out[1:] >= keys_out[:-1])
    n_max = max(len(k) for k in keys_per_rank)
    blk_shift = max(int(np.ceil(np.log2(max(n_max / n_blocks, 1)))), 0)
    blk = np.zeros((world, max(ns, 1)), np.uint32)
    for tv in np.unique(keys_out[tie_out != 0]):
        less_total = int(sum(np.sum(u < tv) for u in okeys))
        eq = np.zeros((world, n_blocks), np.uint32)
        for s, u in enumerate(okeys):
            pos = np.nonzero(u == tv)[0] >> blk_shift
            eq[s] = np.bincount(pos, minlength=n_blocks)[:n_blocks]
        which = np.nonzero((keys_out == tv) & (tie_out != 0))[0]
        targets = np.array([(total * (int(r) + 1)) // world for r in which], np.uint64)
        for rank in range(world):
            out = np.zeros(len(which), np.uint32)
            assert L.b200sort_mgpu_tie_thresholds(world, rank, n_blocks, less_total, eq.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                                  len(which), targets.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))) == 0
            blk[rank, which] = out
    sizes = np.zeros(world, np.int64)
    lo_hi = []
    for rank, u in enumerate(okeys):
        b = (np.arange(len(u)) >> blk_shift).astype(np.uint32)
        dest = np.zeros(len(u), np.int64)
        for r in range(ns):
            dest += ((u > keys_out[r]) | ((u == keys_out[r]) & (b >= blk[rank, r]))).astype(np.int64)
        sizes += np.bincount(dest, minlength=world)
        lo_hi.append((u, dest))
    # destinations are monotonic in the key
    mins = np.full(world, np.iinfo(np.uint64).max, np.uint64)
    maxs = np.zeros(world, np.uint64)
    for u, dest in lo_hi:
        for d in range(world):
            sel = dest == d
            if sel.any():
                mins[d] = min(mins[d], u[sel].min())
                maxs[d] = max(maxs[d], u[sel].max())
    last = None
    for d in range(world):
        if sizes[d]:
            if last is not None:
                assert maxs[last] <= mins[d]
            last = d
    return sizes, total


@pytest.mark.parametrize("world", [2, 3, 5, 8])
@pytest.mark.parametrize("case", ["zero_u8", "two_values_u16", "uniform_u32", "heavy_30pct_u64", "geometric_u16", "distinct_close_u64"])
def test_refined_splitters_balance_any_world(world, case):
    import numpy as np
    n = 30_000
    rngs = [np.random.default_rng(100 * world + r) for r in range(world)]
    if case == "zero_u8":
        ranks = [np.zeros(n + 17 * r, np.uint8) for r in range(world)]
    elif case == "two_values_u16":
        ranks = [np.where(g.random(n) < 0.3, 7, 40000).astype(np.uint16) for g in rngs]
    elif case == "uniform_u32":
        ranks = [g.integers(0, 2**32, size=n, dtype=np.uint32) for g in rngs]
    elif case == "heavy_30pct_u64":
        ranks = [np.where(g.random(n) < 0.3, np.uint64(1) << np.uint64(63), g.integers(0, 2**64, size=n, dtype=np.uint64)) for g in rngs]
    elif case == "geometric_u16":
        ranks = [np.minimum(g.geometric(0.05, size=n), 65535).astype(np.uint16) for g in rngs]
    else:  # distinct keys that differ only in their lowest bits: refinement must go all the way down, no ties
        ranks = [(np.uint64(0xABCD) << np.uint64(48)) + g.permutation(n).astype(np.uint64) + np.uint64(r * n) for r, g in enumerate(rngs)]
    sizes, total = _split_and_count(ranks, world)
    assert sizes.sum() == total
    assert sizes.max() <= 1.125 * total / world + 1, (case, world, sizes.tolist())
