"""Multi-GPU shard path on real GPUs (needs >= 2 devices; skipped otherwise): torchrun with 2 ranks,
NCCL all-reduce / all-gather / all-to-all inside b200sort_mgpu_sort_soa, checked on rank 0 against numpy."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_distributed_sort():
    world = min(torch.cuda.device_count(), 4)
    port = str(29600 + os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", port, str(ROOT / "tests" / "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    for rank in range(world):
        assert f"MGPU_GPU_OK {rank}" in r.stdout


def test_one_rank_communicator_runs_every_multi_gpu_code_path():
    """a world of one on one GPU (NCCL allows it): splitters, counting, the peer-store partition kernel into this
    rank's own landing arrays, the chunked / overlapped exchange with its arrival flags, the forced-plan local sort,
    heavy key values -- everything except real peer traffic.  Runs on the driver's single-GPU box."""
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "mgpu_single.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MGPU_SINGLE_OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
