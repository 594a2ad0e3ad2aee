"""T3: fused communication kernels on >= 2 GPUs (torchrun subprocess, NCCL bootstrap)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_allreduce_adam_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "bench", "dist_check.py"),
                        "--numel", "67108864", "--iters", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(os.path.join(ROOT, "gpurun_out", f"dist_check_{n}.json")))
    assert res["adam_p2p_ranks_identical"]


def test_sharded_multicast_weight_sync_matches_local_merge():
    """K-BC: layer-sharded LoRA merge with multimem.st into every rank's sampler arena == the local merge, bit for bit
    (the check script broadcasts rank 0's adapters first, as the trainers do at start-up)."""
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "bench", "dist_check_wsync.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(os.path.join(ROOT, "gpurun_out", f"wsync_check_{n}.json")))
    assert res["identical_to_local_merge"]


def test_data_parallel_equivalence():
    """T3: N ranks x batch G/N through K-AR == 1 rank accumulating batch G (same seeds), and all ranks end bit-identical."""
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29521", os.path.join(ROOT, "bench", "dist_check_dp.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(os.path.join(ROOT, "gpurun_out", f"dp_equiv_{n}.json")))
    assert res["ok"], res
