"""Worker of tests/test_gpu_multi.py: one rank per GPU under torch.distributed.run."""
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
from _util import GOLDEN, controls_from_dict, load_fixture, solve_keys  # noqa: E402


def check_against_reference_fixture(kind, world, rank, parts, mat, part):
    """tests/golden/decomp<world>_<kind>.b2ls: the unmodified reference executing its decomposed algorithm in serial
    (processor patches as cyclic pairs between diagonal blocks, decompose.as_cyclic_blocks).  Bars as on one GPU:
    iterations +-1, residuals within 1e-9 of the reference on the scale of the initial residual, solution within 1e-9
    when the iteration counts agree.  The residual history of BiCGStab is compared over its first 12 iterations."""
    name = f"decomp{world}_{kind}"
    if not (GOLDEN / f"{name}.b2ls").exists():
        return True
    inp, ref = load_fixture(name)
    blk, offs = decompose.as_cyclic_blocks(parts)
    same = np.array_equal(blk.lower, inp["lower"]) and np.array_equal(blk.diag, inp["diag"]) and \
        np.array_equal(blk.upper_coeffs, inp["upperCoeffs"]) and np.array_equal(blk.source, inp["source"])
    ok = bool(same)
    lo, hi = int(offs[rank]), int(offs[rank + 1])
    worst = 0.0
    for i, text in solve_keys(inp):
        ctl = controls_from_dict(text, recordHistory=1)
        psi, perf = mat.solve(ctl, part.source)
        rperf = ref[f"solve.{i}.perf"]
        good = abs(perf.nIterations - int(rperf[2])) <= 1 and abs(perf.initialResidual - rperf[0]) <= 1e-9 * abs(rperf[0])
        if perf.nIterations == int(rperf[2]):
            rpsi = ref[f"solve.{i}.psi"][lo:hi]
            dpsi = float(np.max(np.abs(psi - rpsi)) / max(np.max(np.abs(rpsi)), 1e-300))
            worst = max(worst, dpsi) if "PBiCGStab" not in text else worst
            sol_tol = max(1e-9, ctl.tolerance) if "PBiCGStab" in text and perf.nIterations > 12 else 1e-9
            good = good and abs(perf.finalResidual - rperf[1]) <= 1e-9 * abs(rperf[0]) and dpsi <= sol_tol and \
                bool(perf.converged) == bool(rperf[3])
        hkey = f"solve.{i}.historyResiduals"
        if hkey in ref:
            h, rh = capi.history(perf), ref[hkey]
            n = min(len(h), len(rh), int(rperf[2]))
            if "PBiCGStab" in text:
                n = min(n, 12)
            good = good and n > 0 and bool(np.all(np.abs(h[:n] - rh[:n]) <= 1e-9 * abs(rperf[0])))
        if not good and rank == 0:
            print(f"{name}: FAIL {text}: iterations {perf.nIterations}/{int(rperf[2])} "
                  f"final {perf.finalResidual:.6e}/{rperf[1]:.6e}", flush=True)
        ok &= bool(good)
    if rank == 0:
        print(f"{name}: {len(list(solve_keys(inp)))} solver configurations vs the reference run as cyclic blocks: "
              f"{'OK' if ok else 'FAIL'} (max solution rel diff {worst:.2e})", flush=True)
    return ok


def dense(sys_):
    a = np.diag(sys_.diag.copy())
    lo_c = sys_.upper_coeffs if sys_.lower_coeffs is None else sys_.lower_coeffs
    a[sys_.upper, sys_.lower] += lo_c
    a[sys_.lower, sys_.upper] += sys_.upper_coeffs
    for itf in sys_.interfaces:                      # cyclic pairs: row gets -bouCoeffs * psi[partner cell]
        np.add.at(a, (itf.face_cells, sys_.interfaces[itf.nbr_patch].face_cells), -itf.bou_coeffs)
    return a


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    capi.init(local, bytes(buf.cpu().numpy().tobytes()), rank, world)

    ok = True
    for kind in ("sym", "asym", "cyc"):
        split = decompose.simple_split(world)
        if kind == "cyc" and split[2] > 1:
            # the periodic direction must stay inside a rank (processorCyclic patches are out of scope)
            split = (split[0] * split[2], split[1], 1)
        nx, ny, nz = 10 * split[0], 8 * split[1], 6 * split[2]
        # "cyc": periodic in z on top of the decomposition -- every rank has processor patches AND a cyclic pair
        glob = cases.cavity_laplacian(nx, ny, nz, coeffs="random") if kind == "sym" else \
            cases.convection_diffusion(nx, ny, nz, dt_coeff=50.0) if kind == "asym" else \
            cases.add_cyclic(cases.cavity_laplacian(nx, ny, nz, coeffs="random"), 2)
        parts, maps = decompose.decompose_system(glob, decompose.box_cell_ranks(nx, ny, nz, split), world)
        part, cells = parts[rank], maps[rank]
        mesh, mat = capi.from_system(part)
        n_coarse = mesh.agglomerate(part.face_weights)      # restrictMap exchange + global stop criterion over NCCL
        mat.set(part.diag, part.upper_coeffs, part.lower_coeffs, [i.bou_coeffs for i in part.interfaces],
                [i.int_coeffs for i in part.interfaces])
        A = dense(glob)
        x = np.cos(0.3 * np.arange(glob.n_cells))

        def gather(local_vec):
            t = torch.zeros(glob.n_cells, dtype=torch.float64, device="cuda")
            t[torch.from_numpy(cells).cuda()] = torch.from_numpy(local_vec).cuda()
            dist.all_reduce(t)
            return t.cpu().numpy()

        y = gather(mat.amul(x[cells]))
        err = np.max(np.abs(y - A @ x)) / np.max(np.abs(A @ x))
        r = gather(mat.residual(x[cells], glob.source[cells]))
        err_r = np.max(np.abs(r - (glob.source - A @ x))) / np.max(np.abs(glob.source))
        ok &= err < 1e-13 and err_r < 1e-13   # vs a dense numpy matvec (different summation order)
        if rank == 0:
            print(f"{kind}: amul rel err {err:.2e} residual rel err {err_r:.2e}", flush=True)

        if kind == "sym":
            # residual history of the decomposed PCG+DIC against the in-process emulation of the reference's
            # decomposed algorithm (tests/_emulated_ranks.py): same iteration count, residuals within 1e-9
            import _emulated_ranks as em

            e_psi, e_perf = em.pcg(parts, "DIC", tolerance=1e-10)
            ctl = capi.controls("PCG", preconditioner="DIC", tolerance=1e-10, relTol=0.0, recordHistory=1)
            psi, perf = mat.solve(ctl, part.source)
            h = capi.history(perf)
            nh = min(len(h), len(e_perf["history"]))
            dh = float(np.max(np.abs(h[:nh] - e_perf["history"][:nh]))) if nh else 0.0
            dpsi = float(np.max(np.abs(psi - e_psi[rank])) / np.max(np.abs(e_psi[rank])))
            good = abs(perf.nIterations - e_perf["nIterations"]) <= 1 and dh <= 1e-9 * e_perf["initialResidual"] and \
                (perf.nIterations != e_perf["nIterations"] or dpsi <= 1e-9)
            ok &= bool(good)
            if rank == 0:
                print(f"sym: decomposed PCG+DIC vs emulated ranks: iterations {perf.nIterations}/{e_perf['nIterations']} "
                      f"max history diff {dh:.2e} solution rel diff {dpsi:.2e} {'OK' if good else 'FAIL'}", flush=True)
        ok &= check_against_reference_fixture(kind, world, rank, parts, mat, part)
        exact = np.linalg.solve(A, glob.source)
        combos = [("PCG", "DIC"), ("PCG", "diagonal")] if kind in ("sym", "cyc") else [("PBiCGStab", "DILU")]
        combos += [("GAMG", "GaussSeidel"), ("GAMG", "DILU" if kind == "asym" else "DIC")]
        # V-cycles as the preconditioner: the coarsest-level Krylov solve is nested inside the outer Krylov loop
        combos.append(("PBiCGStab" if kind == "asym" else "PCG", "GAMG"))
        if kind == "asym":   # Gauss-Seidel alone converges far too slowly on the Laplacian to pin a solution
            combos.append(("smoothSolver", "GaussSeidel"))
        for solver, pre in combos:
            kw = dict(tolerance=1e-13, relTol=0.0, maxIter=2000)
            if solver == "smoothSolver":
                ctl = capi.controls(solver, smoother=pre, nSweeps=4, **kw)
            elif solver == "GAMG":
                ctl = capi.controls(solver, smoother=pre, **kw)
            else:
                ctl = capi.controls(solver, preconditioner=pre, **kw)
            psi, perf = mat.solve(ctl, part.source)
            full = gather(psi)
            e = np.max(np.abs(full - exact)) / np.max(np.abs(exact))
            its = torch.tensor([perf.nIterations], device="cuda")
            lo, hi = its.clone(), its.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            good = e < 1e-9 and int(lo) == int(hi) and perf.converged
            ok &= bool(good)
            if rank == 0:
                print(f"{kind}: {solver}+{pre}: iterations {perf.nIterations} solution rel err {e:.2e} "
                      f"converged {perf.converged} {'OK' if good else 'FAIL'}", flush=True)
        mat.close()
        mesh.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK" if int(flag) else "MGPU_FAIL", flush=True)
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
