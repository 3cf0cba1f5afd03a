"""GAMG + GaussSeidel V-cycles on the N^3 cavity, decomposed `simple` over the ranks of a torchrun launch (or one GPU).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/gamg_strong.py 384 [cycles]

With B200LS_VPROF=<prefix> the library appends a per-phase timing of the V-cycles to <prefix>.rank<r>.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from _pkg import load_pkg  # noqa: E402

load_pkg()
from b200ls import capi, cases, decompose  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
uid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())
capi.init(local_rank, uid, rank, world)

if world == 1:
    part, split = cases.cavity_laplacian(n, n, n), (1, 1, 1)
else:
    split = decompose.simple_split(world)
    part = decompose.cavity_subdomain(n, n, n, split, rank)
mesh, mat = capi.from_system(part)
t0 = time.perf_counter()
mesh.agglomerate(part.face_weights)
t_agg = time.perf_counter() - t0
mat.set(part.diag, part.upper_coeffs, None, [i.bou_coeffs for i in part.interfaces],
        [i.int_coeffs for i in part.interfaces])
smoother = os.environ.get("SMOOTHER", "GaussSeidel")
ctl = capi.controls("GAMG", smoother=smoother, tolerance=1e-6, relTol=0.01, minIter=cycles, maxIter=cycles)
d_src = torch.from_numpy(part.source).cuda()
d_psi = torch.zeros(part.n_cells, dtype=torch.float64, device="cuda")
for rep in range(3):
    d_psi.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    perf = mat.solve_dev(ctl, d_psi.data_ptr(), d_src.data_ptr())
    t = torch.tensor([perf.solveMs / max(1, perf.nIterations)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sizes = [int(mesh.get_i32(12, k)[0]) for k in range(mesh.n_levels)]
        print(f"{n}^3 split {split} {smoother}: rep {rep} ms/cycle {float(t[0]):.3f} setupMs {perf.setupMs:.2f} "
              f"launches/cycle {perf.kernelLaunches / max(1, perf.nIterations):.0f} residual {perf.finalResidual:.6e} "
              f"levels {len(sizes)} agglomeration {t_agg:.1f}s", flush=True)
if rank == 0:
    print("level cells (rank 0):", sizes)
mat.close()
mesh.close()
if world > 1:
    dist.destroy_process_group()
