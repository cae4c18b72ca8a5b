#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 sort (BASELINE.json metric: Gpairs/s, key+payload sorted).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one complete sort of one batch of synthetic records.
  N = 1 : BASELINE.json configs[1]: uint64 key + uint64 payload, 1e9 uniform-random records, ascending.
  N > 1 : configs[4] (weak scaling): every rank holds 1e9 such records; one step = sampled top histogram ->
          NCCL all-reduce -> splitters -> exact destination counts -> NCCL all-gather -> ONE scatter pass that
          writes every record straight into its destination GPU's memory over NVLink -> local sort.
`value` is device-resident throughput (records of all ranks / max-over-ranks CUDA-event time of the
sort itself; the untimed copy that restores the unsorted input between steps is outside the events).
`e2e` is the same sort through the public API with HOST (pinned) buffers, H2D + D2H inside the timing.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, else the oracle port).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_RECORDS = 1_000_000_000          # per GPU
RECORD_BYTES = 16
KEY_BYTES = 8
METRIC = "Gpairs/s (key+payload) sorted"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", dest="n", type=int, default=N_RECORDS, help="records per GPU (development override)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=1 << 27)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# CPU baseline (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------
def cpu_sort_sample(sample_n: int, seed: int = 12345):
    """Sorts a bounded sample of the same workload on the host with the reference's own implementation
    (oracle/_ref: radixSort.hpp compiled as is, single-threaded like the reference) or, where that
    cannot run, the oracle port.  Returns (seconds, kind, sample description, cores)."""
    import oracle_lib as O

    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 2**64, size=sample_n, dtype=np.uint64)
    pay = np.arange(sample_n, dtype=np.uint64)
    if O.ref_available():
        kind, fn = "reference", O.ref_sort_soa
    else:
        kind, fn = "port", O.port_sort_soa
        sample_n = min(sample_n, 1 << 23)
        keys, pay = keys[:sample_n].copy(), pay[:sample_n].copy()
    t0 = time.perf_counter()
    fn(keys, [pay], True)
    dt = time.perf_counter() - t0
    assert bool(np.all(keys[:-1] <= keys[1:]))
    return dt, kind, f"{sample_n} uniform uint64 keys + uint64 payloads (same generator family as the GPU run), 1 thread", 1, sample_n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    sample_n = args.cpu_sample
    kind = "reference"
    sample = ""
    for i in range(args.warmup + args.steps):
        dt, kind, sample, cores, used_n = cpu_sort_sample(sample_n, seed=12345 + i)
        if i >= args.warmup:
            vals.append(used_n / dt * 1e-9)
        if dt > 40:  # keep the whole run within minutes on a slow host
            sample_n = max(sample_n // 2, 1 << 20)
    v = statistics.mean(vals)
    ms = 1e3 * (1.0 / v) * 1e-9 * N_RECORDS
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Gpairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "uint64 key + uint64 payload, uniform random, ascending (BASELINE.json configs[1]); "
                               "each step sorts a bounded sample on the host", "records_per_step": sample_n},
        "cpu_baseline": {"value": v, "unit": "Gpairs/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in Path(self.tmp.name).read_text().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        import datetime
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                if t_begin is not None:
                    ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t_begin - 0.05 or ts > t_end + 0.05:
                        continue
                sm.append(float(r[2])); mx.append(float(r[3]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for name, val in zip(names, r[6:10]):
                    if val.strip().lower() == "active":
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(kernel: str):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture, scaled to
    this run's launch size; None if no capture is committed."""
    p = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if not p.exists():
        return None
    try:
        j = json.loads(p.read_text())
        return j.get(kernel)
    except Exception:
        return None


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import simd_radix_sort_b200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sort has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    comm = None
    if multi:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        import ctypes
        uid = [None]
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            rc = S.lib().b200sort_mgpu_unique_id(buf)
            if rc:
                raise S.B200SortError(rc, S.lib().b200sort_last_error().decode())
            uid[0] = bytes(buf)
        dist.broadcast_object_list(uid, src=0)
        comm = ctypes.c_void_p()
        idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(uid[0])
        rc = S.lib().b200sort_mgpu_comm_create(ctypes.byref(comm), world, rank, idbuf)
        if rc:
            raise S.B200SortError(rc, S.lib().b200sort_last_error().decode())

    n = args.n
    cap = n if not multi else int(n * 1.125) + 4096
    gen = torch.Generator(device=dev)
    gen.manual_seed(12345 + rank)
    keys0 = torch.randint(-(2**63), 2**63 - 1, (n,), dtype=torch.int64, device=dev, generator=gen).view(torch.uint64)
    pay0 = (torch.arange(n, dtype=torch.int64, device=dev) + rank * n).view(torch.uint64)
    keys = torch.empty(cap, dtype=torch.uint64, device=dev)
    pay = torch.empty(cap, dtype=torch.uint64, device=dev)
    key_sum = int(keys0.view(torch.int64).sum().item())  # wraps mod 2^64: a permutation-invariant checksum

    def restore():
        keys[:n].copy_(keys0)
        pay[:n].copy_(pay0)

    out_n = [n]

    def do_sort():
        if not multi:
            S.sort(n, keys, pay, up=True)
        else:
            import ctypes
            ptrs = (ctypes.c_void_p * 1)(pay.data_ptr())
            sizes = (ctypes.c_uint32 * 1)(8)
            got = ctypes.c_int64(0)
            rc = S.lib().b200sort_mgpu_sort_soa(comm, keys.data_ptr(), S.KEY_TYPES["uint64"], n, cap, 1, 1, ptrs, sizes,
                                                ctypes.byref(got), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc:
                raise S.B200SortError(rc, S.lib().b200sort_last_error().decode())
            out_n[0] = got.value

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a moment to start; only samples inside the timed region are kept
    # ---- warm-up (also the correctness gate: a fast wrong sort is not a result) ----
    for w in range(args.warmup):
        restore()
        do_sort()
    torch.cuda.synchronize()
    m = out_n[0]
    ks = keys[:m].view(torch.int64)
    # ascending as unsigned <=> ascending after flipping the sign bit as signed
    flipped = ks ^ (-(2**63))
    assert bool((flipped[1:] >= flipped[:-1]).all().item()), "result is not sorted"
    if not multi:
        assert int(ks.sum().item()) == key_sum, "key multiset changed"
        probe = torch.randint(0, n, (1 << 20,), device=dev)
        src = pay[:n].view(torch.int64)[probe]
        assert bool((keys0.view(torch.int64)[src] == ks[probe]).all().item()), "payload did not follow its key"
    else:
        tot = torch.tensor([m], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        assert int(tot.item()) == n * world, "records were lost in the exchange"
        # rank boundaries: my last key <= next rank's first key
        edge = torch.stack([flipped[0], flipped[-1]]) if m > 0 else torch.zeros(2, dtype=torch.int64, device=dev)
        edges = [torch.zeros_like(edge) for _ in range(world)]
        dist.all_gather(edges, edge)
        for r in range(world - 1):
            assert int(edges[r][1].item()) <= int(edges[r + 1][0].item()), "rank ranges overlap"
    del flipped, ks

    # ---- timed steps ----
    S.set_option("profile", 1)
    step_ms = []
    prof = []
    launches0 = S.launch_count()
    barrier()
    t_begin = time.time()
    for k in range(args.steps):
        restore()
        if multi:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        do_sort()
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        prof.append(S.last_profile())
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    launches = S.launch_count() - launches0
    S.set_option("profile", 0)
    stats = S.last_stats()

    total_ms = sum(step_ms)
    if multi:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = (n * world) / (ms_per_step * 1e-3) * 1e-9

    # ---- end-to-end through the public API with host buffers ----
    e2e = None
    h2d = d2h = n * RECORD_BYTES
    try:
        if args.e2e_steps <= 0:
            raise RuntimeError("skipped (--e2e-steps 0)")
        hk = torch.empty(n, dtype=torch.uint64, pin_memory=True)
        hp = torch.empty(n, dtype=torch.uint64, pin_memory=True)
        e2e_ms = []
        for k in range(args.e2e_steps + 1):  # the first one is a warm-up (staging buffers get allocated): not timed
            hk.copy_(keys0); hp.copy_(pay0)
            torch.cuda.synchronize()
            if multi:
                dist.barrier()
            t0 = time.perf_counter()
            if not multi:
                S.sort(n, hk.numpy(), hp.numpy(), up=True)  # host pointers: H2D, sort, D2H inside the call
            else:
                keys[:n].copy_(hk, non_blocking=True); pay[:n].copy_(hp, non_blocking=True)
                do_sort()
                m = out_n[0]
                hk[:min(m, n)].copy_(keys[:min(m, n)], non_blocking=True); hp[:min(m, n)].copy_(pay[:min(m, n)], non_blocking=True)
                torch.cuda.synchronize()
            if k > 0:
                e2e_ms.append((time.perf_counter() - t0) * 1e3)
        e2e_t = statistics.mean(e2e_ms)
        if multi:
            t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_t = float(t.item())
        hkn = hk.numpy()
        assert multi or bool(np.all(hkn[:-1:4097] <= hkn[1::4097])), "host result not sorted"
        e2e = {"value": (n * world) / (e2e_t * 1e-3) * 1e-9, "unit": "Gpairs/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_t, "steps": args.e2e_steps, "warmup": 1}
        del hk, hp
    except Exception as ex:  # e.g. the box cannot pin 16 GB
        e2e = {"value": None, "unit": "Gpairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "error": str(ex)[:200]}

    if rank != 0:
        if multi:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the digit scatter pass) ----
    peak, peak_src = measured_peak_gbs()
    # scatter launches of skipped digit positions return at once; only executed passes count
    sweep_all = [ms for step in prof for kind, ms in step if kind == "sweep"]
    n_exec = max(int(stats["passes_planned"]), 1)
    sweep_ms = sorted(sweep_all, reverse=True)[: n_exec * args.steps] if stats["algo"] == 2 else sweep_all
    per_kind = {}
    for step in prof:
        for kind, ms in step:
            per_kind[kind] = per_kind.get(kind, 0.0) + ms / args.steps
    n_sweep_launch = stats["num"]  # records one scatter launch moves (the local sort's size on this rank)
    alg_bytes = 2 * n_sweep_launch * RECORD_BYTES
    avg_sweep = statistics.mean(sweep_ms) if sweep_ms else None
    achieved = alg_bytes / (avg_sweep * 1e-3) * 1e-9 if avg_sweep else None
    traffic = ncu_traffic_per_launch("onesweep_kernel")
    roofline = {"bound": "hbm", "kernel": "onesweep_kernel (digit scatter pass)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "peak_source": peak_src,
                "traffic": traffic * (n_sweep_launch / 67108864) if traffic else None,
                "traffic_source": "profiles/dominant_kernel_traffic.json (ncu --set full at 2^26 records, scaled linearly)" if traffic else None,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_sweep,
                "sweep_launches_per_step": len(sweep_ms) / args.steps,
                "ms_per_step_by_kernel": per_kind,
                "whole_sort": {"algorithmic_bytes": stats["algorithmic_bytes"],
                               "achieved_gbs": stats["algorithmic_bytes"] / (ms_per_step * 1e-3) * 1e-9 if not multi else None,
                               "floor_2NR_gbs": 2 * n * world * RECORD_BYTES / (ms_per_step * 1e-3) * 1e-9}}

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only) ----
    cpu = None
    if not multi and args.cpu_sample > 0:
        dt, kind, sample, cores, used_n = cpu_sort_sample(args.cpu_sample)
        cpu = {"value": used_n / dt * 1e-9, "unit": "Gpairs/s", "cores": cores, "kind": kind, "sample": sample}

    exchange = None
    if multi:
        S.lib().b200sort_mgpu_used_p2p.argtypes = [ctypes.c_void_p]
        exchange = ("scatter pass writes straight into peer memory (cudaIpc-mapped workspaces)"
                    if S.lib().b200sort_mgpu_used_p2p(comm) else "local scatter pass + ncclSend/ncclRecv")
    line = {
        "metric": METRIC, "value": value, "unit": "Gpairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": ("uint64 key + uint64 payload, 1e9 uniform-random records on 1 B200 (BASELINE.json configs[1])"
                                if not multi else
                                f"uint64 key + uint64 payload, {world}e9 records sharded over {world} B200, exchanged over NVLink "
                                "(BASELINE.json configs[4], 1e9 per GPU)"),
                   "records_per_gpu": n, "record_bytes": RECORD_BYTES, "ascending": True,
                   "l2": "inputs (16 GB per GPU) are larger than L2; no flush needed",
                   "algo": {1: "LSD one-sweep", 2: "hybrid MSB"}.get(stats["algo"], "?"),
                   "scatter_passes": stats["passes_planned"], "hist_sweeps": stats["hist_sweeps"],
                   **({"exchange": exchange} if multi else {})},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
