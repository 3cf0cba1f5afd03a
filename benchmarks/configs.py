#!/usr/bin/env python3
"""Runs one BASELINE.json configuration and prints one JSON line (used to fill profiles/rNN_configs.jsonl).

    python benchmarks/configs.py gamg256                       # configs[2]: 256^3, GAMG + GaussSeidel, 1 GPU
    torchrun --nproc-per-node 8 benchmarks/configs.py cavity384 --solver PCG|GAMG   # configs[3]: 384^3 / N GPUs
    [torchrun ...] benchmarks/configs.py convdiff12m           # configs[4]: asymmetric 3538^2 (12.5 M cells), PBiCGStab+DILU
    python benchmarks/configs.py cavity128 --solver GAMG       # extra

Time is the library's CUDA-event time of the solve (perf.solveMs + perf.setupMs), max over ranks; --cpu also times
the unmodified reference (oracle/_ref, one core) on the same single-rank system with a bounded iteration count.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
from _pkg import load_pkg  # noqa: E402

load_pkg()
from b200ls import capi, cases, decompose  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cavity128", "gamg256", "cavity384", "convdiff12m"])
    ap.add_argument("--solver", default=None)
    ap.add_argument("--smoother", default="GaussSeidel")
    ap.add_argument("--tolerance", type=float, default=1e-6)
    ap.add_argument("--relTol", type=float, default=0.01)
    ap.add_argument("--maxIter", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--n", type=int, default=0, help="override the side length")
    args = ap.parse_args()

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    uid = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
    capi.init(local, uid, rank, world)

    t0 = time.time()
    if args.config in ("cavity128", "gamg256", "cavity384"):
        n = args.n or {"cavity128": 128, "gamg256": 256, "cavity384": 384}[args.config]
        solver = args.solver or ("GAMG" if args.config == "gamg256" else "PCG")
        split = decompose.simple_split(world)
        if world == 1:
            sys_ = cases.cavity_laplacian(n, n, n)
        else:
            sys_ = decompose.cavity_subdomain(n, n, n, split, rank)
        desc = f"cavity {n}^3 p-equation" + (f" decomposed simple {split}" if world > 1 else "")
        n_total = n ** 3
    else:
        n = args.n or 3538
        solver = args.solver or "PBiCGStab"
        split = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (4, 2, 1)}[world]
        n = (n // (split[0] * split[1])) * split[0] * split[1] if world > 1 else n
        if world == 1:
            sys_ = cases.convection_diffusion(n, n, 1, dt_coeff=50.0)
        else:
            glob = cases.convection_diffusion(n, n, 1, dt_coeff=50.0)
            parts, _ = decompose.decompose_system(glob, decompose.box_cell_ranks(n, n, 1, split), world)
            sys_ = parts[rank]
            del glob, parts
        desc = f"asymmetric convection-diffusion {n}x{n} (pitzDaily-sized stand-in)" + \
            (f" decomposed simple {split}" if world > 1 else "")
        n_total = n * n
    t_gen = time.time() - t0

    t0 = time.time()
    mesh, mat = capi.from_system(sys_)
    t_mesh = time.time() - t0
    t_agg = 0.0
    if solver == "GAMG":
        t0 = time.time()
        n_coarse = mesh.agglomerate(sys_.face_weights)
        t_agg = time.time() - t0
        mat.set(sys_.diag, sys_.upper_coeffs, sys_.lower_coeffs, [i.bou_coeffs for i in sys_.interfaces],
                [i.int_coeffs for i in sys_.interfaces])
    kw = dict(tolerance=args.tolerance, relTol=args.relTol, maxIter=args.maxIter)
    if solver == "GAMG":
        ctl = capi.controls("GAMG", smoother=args.smoother, **kw)
        name = f"GAMG+{args.smoother}"
    else:
        pre = "DIC" if sys_.symmetric else "DILU"
        ctl = capi.controls(solver, preconditioner=pre, **kw)
        name = f"{solver}+{pre}"

    best = None
    for _ in range(args.reps):
        psi, perf = mat.solve(ctl, sys_.source)
        ms = perf.solveMs + perf.setupMs
        if best is None or ms < best[0]:
            best = (ms, perf.nIterations, perf.initialResidual, perf.finalResidual, perf.kernelLaunches, perf.setupMs)
    ms, its, ini, fin, launches, setup_ms = best
    if world > 1:
        import torch
        import torch.distributed as dist

        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    line = {
        "config": args.config, "workload": desc, "solver": name, "n_gpus": world, "n_cells": n_total,
        "tolerance": args.tolerance, "relTol": args.relTol, "iterations": its, "initialResidual": ini,
        "finalResidual": fin, "solve_ms": ms, "ms_per_iteration": ms / max(its, 1),
        "cell_iterations_per_s": n_total * its / (ms * 1e-3), "setup_ms_in_solve": setup_ms,
        "kernel_launches": int(launches), "host_mesh_analysis_s": round(t_mesh, 2), "host_agglomeration_s": round(t_agg, 2),
        "case_generation_s": round(t_gen, 2),
    }
    if args.cpu and world == 1 and rank == 0:
        import bench

        bounded = 20 if solver != "GAMG" else its
        from b200ls import ldu_io
        import subprocess
        import tempfile

        e = cases.to_entries(sys_)
        d = f"solver {solver}; tolerance {args.tolerance}; relTol {args.relTol}; maxIter {min(bounded, its)};"
        d += f" smoother {args.smoother};" if solver == "GAMG" else f" preconditioner {pre};"
        e["solve.0.dict"] = d
        with tempfile.TemporaryDirectory() as td:
            ldu_io.write(f"{td}/in.b2ls", e)
            r = subprocess.run([str(ROOT / "oracle/_ref/ref_harness"), f"{td}/in.b2ls", f"{td}/out.b2ls", f"{td}/case"],
                               env=bench.ref_env(), capture_output=True, text=True)
            if r.returncode == 0:
                out = ldu_io.read(f"{td}/out.b2ls")["solve.0.perf"]
                line["reference_cpu_1core"] = {"iterations": int(out[2]), "seconds": float(out[5]),
                                               "ms_per_iteration": 1e3 * float(out[5]) / max(int(out[2]), 1),
                                               "finalResidual": float(out[1])}
            else:
                line["reference_cpu_1core"] = {"error": r.stderr[-300:]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
