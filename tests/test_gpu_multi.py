"""GPU tier, N > 1: the real NCCL path.  Two (or more) processes, one per GPU, each on its contiguous shard;
the concatenated results and the all-reduced counts must equal the single-process oracle on the whole input
(v3, sci-RNA-seq3 and ATAC-like layouts, two chunks per rank).  Skipped on a box with one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["host-built-tables", "device-built-tables", "device-built-slot-order"])
def test_two_gpus_reproduce_the_single_process_oracle(mode):
    """mode: small whitelists are hashed on the host; PCL_DEVICE_BUILD_MIN=1 sends them through the device build, whose
    placement by rank must give every GPU the same slots (the all-reduce sums the counts slot by slot) -- with the
    occupied slots in ascending order (default) or the whole slot array (PCL_REDUCE_SLOTS=1)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    env = dict(os.environ)
    if mode != "host-built-tables":
        env["PCL_DEVICE_BUILD_MIN"] = "1"
    if mode == "device-built-slot-order":
        env["PCL_REDUCE_SLOTS"] = "1"
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "MULTI_GPU_PARITY_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
