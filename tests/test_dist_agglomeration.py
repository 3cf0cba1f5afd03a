"""N>1 host logic on CPU: world_size 2 and 4 `gloo` process groups run the multi-rank GAMG agglomeration through
the C-ABI (b200ls_set_host_comm + b200ls_agglomerate) and check the coarse processor interfaces for consistency
(processorGAMGInterface.C:53-140, GAMGAgglomeration.C:205-230)."""
import subprocess
import sys
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent


@pytest.mark.parametrize("n", [2, 4])
def test_multi_rank_agglomeration_gloo(n):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + n), str(HERE / "_dist_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
