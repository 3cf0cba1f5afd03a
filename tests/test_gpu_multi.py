"""Multi-GPU path: one subdomain per GPU, NCCL send/recv halos and NCCL allreduce dot products.  Needs >= 2 GPUs
(skipped on a single-GPU box).  The decomposed operator must equal the global operator and every solver must
converge to the global solution with identical iteration counts on all ranks."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

HERE = Path(__file__).resolve().parent


@pytest.mark.parametrize("n", [2, 4, 8])
def test_decomposed_solves(n):
    import torch

    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29510 + n), str(HERE / "_mgpu_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
