"""z-slabs over NCCL on real GPUs (needs >= 2 visible devices; skipped otherwise): every rank's slab of a run with the halo
exchange overlapped with the interior equals a single-GPU run of the whole box bit for bit -- periodic + LES, V60 with every
feature, the fused pressure-gradient drive (rho planes in the halo), the legacy solver with FD-LES -- and the two-way particle
coupling on slabs (replicated particles, owner computes, packed all-reduce) matches the single-GPU coupling
(scripts/check_slabs.py, launched with torchrun)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_nccl_ranks_reproduce_the_single_gpu_run():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "scripts", "check_slabs.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0 and "[check_slabs] ALL OK" in r.stdout, tail
