"""Worker of tests/test_dist_agglomeration.py: world_size-N gloo processes on CPU exercising the host-side
multi-rank agglomeration (restrictMap exchange over processor patches, global stop criterion, coarse interfaces)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _pkg import load_pkg  # noqa: E402

load_pkg()
from b200ls import capi, cases, decompose  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")

    def exchange(nbr, send):
        recv = [torch.empty(len(s), dtype=torch.int32) for s in send]
        reqs = [dist.irecv(r, src=n) for r, n in zip(recv, nbr) if r.numel()]
        reqs += [dist.isend(torch.from_numpy(s.astype(np.int32)), dst=n) for s, n in zip(send, nbr) if len(s)]
        for q in reqs:
            q.wait()
        return [r.numpy() for r in recv]

    def total(v):
        t = torch.tensor([v], dtype=torch.int64)
        dist.all_reduce(t)
        return int(t)

    capi.set_host_comm(rank, world, exchange, total)
    split = decompose.simple_split(world)
    nx, ny, nz = 12 * split[0], 10 * split[1], 8 * split[2]
    glob = cases.cavity_laplacian(nx, ny, nz, coeffs="random")
    parts, maps = decompose.decompose_system(glob, decompose.box_cell_ranks(nx, ny, nz, split), world)
    part = parts[rank]
    mesh = capi.Mesh(part.n_cells, part.lower, part.upper, part.interfaces)
    n_coarse = mesh.agglomerate(part.face_weights)
    ok = n_coarse >= 3

    # every rank created the same number of levels and the global stop criterion holds
    t = torch.tensor([n_coarse, -n_coarse], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok &= int(t[0]) == n_coarse and int(-t[1]) == n_coarse
    coarsest_cells = total(int(mesh.get_i32(capi.LEVEL_SIZES, n_coarse)[0]))
    ok &= coarsest_cells >= 10 * world

    n_if = len(part.interfaces)
    for lev in range(n_coarse):
        restrict = mesh.get_i32(capi.RESTRICT_ADDRESSING, lev)
        for i in range(n_if):
            fc = mesh.get_iface_i32(0, lev, i)
            pr = mesh.get_iface_i32(1, lev, i)
            cfc = mesh.get_iface_i32(0, lev + 1, i)
            # coarse faceCells are the coarse cells of the fine patch cells, first-seen order
            ok &= bool(np.array_equal(cfc[pr], restrict[fc]))
            ok &= pr.max() + 1 == cfc.size and bool(np.all(np.diff(np.unique(pr, return_index=True)[1]) > 0))
            # both sides number the coarse patch faces identically
            other = exchange([part.interfaces[i].neighb_rank], [pr])[0]
            ok &= bool(np.array_equal(other, pr))
            # ... and pair the same coarse cells: my coarse cell of coarse face k <-> neighbour's coarse cell
            mine_pairs = np.stack([cfc[pr], exchange([part.interfaces[i].neighb_rank], [cfc[pr]])[0]], axis=1)
            # a coarse patch face is one distinct (mine, theirs) pair
            ok &= len({tuple(r) for r in mine_pairs.tolist()}) == cfc.size
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_OK" if int(flag) else "DIST_FAIL", "levels", n_coarse, "coarsest cells", coarsest_cells, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
