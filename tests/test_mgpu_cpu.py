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
