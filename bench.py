#!/usr/bin/env python3
"""bench.py -- pressure-solve throughput of libb200ls on BASELINE.json's configuration.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU solver (oracle/_ref)

A "step" is one pressure solve: PCG + DIC, fixed 50 iterations (tolerance 0, relTol 0, maxIter 50 -- the timing
mode of SURVEY.md 8(d)) on the synthetic lid-driven-cavity p-equation of the named size.
  N = 1 : BASELINE.json configs[1]: 128^3 cells (2,097,152) on one B200.
  N > 1 : configs[3] style: a (N*128)x128x128 ... block decomposed `simple` into N subdomains of 128^3 cells, one
          per GPU, processor-patch halos over NCCL send/recv and gSumProd/gSumMag over NCCL allreduce (weak scaling).
metric  = cell-iterations/s = nCells_total * iterations / time.
value   : matrix, psi and source resident in HBM when the timed region starts (b200ls_solve_dev).
e2e     : the same solve through the host-pointer C-ABI calls a plugin makes (b200ls_matrix_set + b200ls_solve):
          coefficient/psi/source H2D and psi D2H inside the timed region, from pinned host memory.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "tests"))

N_SIDE = 128
ITERS = 50
METRIC = "pressure-solve throughput (PCG+DIC, cavity p-equation)"
UNIT = "cell-iterations/s"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ref_env():
    return dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM",
                WM_PROJECT_VERSION="dev")


def run_reference_sample(sys_, iters):
    """Time the UNMODIFIED reference PCG+DIC (oracle/_ref/ref_harness, serial Pstream/dummy) on the same matrix.
    Returns (cell-iterations/s, seconds, iterations)."""
    from b200ls import cases, ldu_io

    harness = ROOT / "oracle/_ref/ref_harness"
    if not harness.exists():
        return None      # callers fall back to the C port of the oracle
    e = cases.to_entries(sys_)
    e.pop("faceWeights", None)
    e["solve.0.dict"] = f"solver PCG; preconditioner DIC; tolerance 0; relTol 0; maxIter {iters};"
    with tempfile.TemporaryDirectory() as td:
        ldu_io.write(f"{td}/in.b2ls", e)
        r = subprocess.run([str(harness), f"{td}/in.b2ls", f"{td}/out.b2ls", f"{td}/case"], env=ref_env(),
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference harness failed: " + r.stderr[-500:])
        out = ldu_io.read(f"{td}/out.b2ls")
    perf = out["solve.0.perf"]
    secs, its = float(perf[5]), int(perf[2])
    return sys_.n_cells * its / secs, secs, its


def run_reference_all_cores(nx, ny, nz, iters_full, target_s=4.0):
    """The reference with every host core it can use.  The image has no MPI, so `mpirun -np P` cannot be run; what CAN
    be measured is its upper bound: P concurrent, independent, serial reference processes, each running PCG+DIC on one
    z-slab of the nx*ny*nz cavity matrix (nx*ny*(nz/P) cells -- the rank-local work of a `simple (1 1 P)` decomposition
    without any halo exchange or global reduction), all competing for the same memory system.  Aggregate throughput =
    total cells * iterations / slowest process.  P = largest power of two <= min(cores, 64) that divides nz.
    Returns (cell-iterations/s, P, description) or None when oracle/_ref is absent."""
    from b200ls import cases, ldu_io

    harness = ROOT / "oracle/_ref/ref_harness"
    if not harness.exists():
        return None
    cores = os.cpu_count() or 1
    P = 1
    while P * 2 <= min(cores, 64) and nz % (P * 2) == 0:
        P *= 2
    slab = cases.cavity_laplacian(nx, ny, nz // P)
    # the same work for every process; enough iterations that start-up jitter does not matter
    iters = int(min(2000, max(iters_full, iters_full * P // 2)))
    e = cases.to_entries(slab)
    e.pop("faceWeights", None)
    e["solve.0.dict"] = f"solver PCG; preconditioner DIC; tolerance 0; relTol 0; maxIter {iters};"
    with tempfile.TemporaryDirectory() as td:
        ldu_io.write(f"{td}/in.b2ls", e)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([str(harness), f"{td}/in.b2ls", f"{td}/out{k}.b2ls", f"{td}/case{k}"], env=ref_env(),
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for k in range(P)]
        rcs = [q.wait() for q in procs]
        wall = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("reference harness failed")
        secs = [float(ldu_io.read(f"{td}/out{k}.b2ls")["solve.0.perf"][5]) for k in range(P)]
    value = slab.n_cells * P * iters / max(secs)
    desc = (f"{P} concurrent serial reference processes (unmodified lduMatrix::solver, oracle/_ref), each {iters} PCG+DIC "
            f"iterations on a {nx}x{ny}x{nz // P} slab of the {nx}x{ny}x{nz} matrix, no halo exchange: UPPER BOUND of "
            f"`mpirun -np {P}` on this host ({cores} cores; the image has no MPI); slowest process {max(secs):.2f} s, "
            f"wall {wall:.1f} s")
    return value, P, desc


def run_oracle_port_sample(sys_, iters):
    """Fallback when oracle/_ref is absent: time the plain-C restatement (oracle/ldu_oracle.c, one core)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import ldu_oracle as orc

    S = orc.System(sys_)
    t0 = time.perf_counter()
    _, perf = orc.solve(S, "PCG", orc.controls("DIC", tolerance=0.0, relTol=0.0, maxIter=iters), sys_.source)
    secs = time.perf_counter() - t0
    return sys_.n_cells * perf["nIterations"] / secs, secs, perf["nIterations"]


def bench_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the same global workload as our arm
    (the undecomposed (128*px)x(128*py)x(128*pz) cavity matrix; iterations per step bounded to ceil(50/N) so the run
    ends within minutes).  The image has no MPI, so the reference runs as one serial process (Pstream/dummy); under
    torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from _pkg import load_pkg

    load_pkg()
    from b200ls import cases, decompose, ldu_io

    px, py, pz = decompose.simple_split(args.gpus)
    sys_ = cases.cavity_laplacian(N_SIDE * px, N_SIDE * py, N_SIDE * pz)
    iters = max(5, -(-ITERS // args.gpus))
    harness = ROOT / "oracle/_ref/ref_harness"
    kind = "reference" if harness.exists() else "port"
    secs_all = []
    with tempfile.TemporaryDirectory() as td:
        if kind == "reference":
            e = cases.to_entries(sys_)
            e.pop("faceWeights", None)
            e["solve.0.dict"] = f"solver PCG; preconditioner DIC; tolerance 0; relTol 0; maxIter {iters};"
            ldu_io.write(f"{td}/in.b2ls", e)
        for i in range(args.warmup + args.steps):
            if kind == "reference":
                r = subprocess.run([str(harness), f"{td}/in.b2ls", f"{td}/out.b2ls", f"{td}/case"], env=ref_env(),
                                   capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError("reference harness failed: " + r.stderr[-500:])
                perf = ldu_io.read(f"{td}/out.b2ls")["solve.0.perf"]
                secs, its = float(perf[5]), int(perf[2])
            else:
                _, secs, its = run_oracle_port_sample(sys_, iters)
            if i >= args.warmup:
                secs_all.append(secs)
    single = sys_.n_cells * iters * len(secs_all) / sum(secs_all)
    value, cores, sample = single, 1, None
    if kind == "reference":
        try:
            multi = run_reference_all_cores(N_SIDE * px, N_SIDE * py, N_SIDE * pz, ITERS)
            if multi and multi[0] > single:
                value, cores, sample = multi
        except Exception:
            pass
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sys_.n_cells * iters / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cavity {N_SIDE * px}x{N_SIDE * py}x{N_SIDE * pz} p-equation (undecomposed), PCG+DIC, "
                               f"{iters} iterations per solve (bounded sample of the {ITERS}-iteration step)",
                   "n_cells": sys_.n_cells, "iterations_per_step": iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample or (
                             f"{args.steps} x {iters} PCG+DIC iterations on the {sys_.n_cells}-cell matrix, " +
                             ("unmodified reference lduMatrix::solver (serial Pstream/dummy; the image has no MPI)"
                              if kind == "reference" else "C restatement oracle/ldu_oracle.c (oracle/_ref absent)"))},
        "single_core": {"value": single, "unit": UNIT,
                        "sample": f"{args.steps} x {iters} PCG+DIC iterations on the whole {sys_.n_cells}-cell matrix, "
                                  "one serial reference process"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def bench_ours(args):
    import torch
    import torch.distributed as dist

    from _pkg import load_pkg

    load_pkg()
    from b200ls import capi, cases, decompose

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch N>1 with torch.distributed.run (one rank per GPU)")
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

    # workload: this rank's subdomain
    if world == 1:
        sys_ = cases.cavity_laplacian(N_SIDE, N_SIDE, N_SIDE)
        workload = f"cavity {N_SIDE}^3 p-equation (BASELINE configs[1]), PCG+DIC, {ITERS} iterations per solve"
    else:
        px, py, pz = decompose.simple_split(world)
        sys_ = decompose.cavity_subdomain(N_SIDE * px, N_SIDE * py, N_SIDE * pz, (px, py, pz), rank)
        workload = (f"cavity {N_SIDE * px}x{N_SIDE * py}x{N_SIDE * pz} p-equation decomposed simple ({px} {py} {pz}), "
                    f"{N_SIDE}^3 cells per GPU, PCG+DIC, {ITERS} iterations per solve")
    n_local = sys_.n_cells
    n_total = n_local * world
    n_faces = sys_.n_faces

    mesh, mat = capi.from_system(sys_)
    ctl = capi.controls("PCG", "DIC", tolerance=0.0, relTol=0.0, maxIter=ITERS)

    # pinned host copies (e2e) and device-resident inputs (value)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    _, h_diag = pinned(sys_.diag)
    _, h_upper = pinned(sys_.upper_coeffs)
    _, h_source = pinned(sys_.source)
    t_psi, h_psi = pinned(np.zeros(n_local))
    bou = [i.bou_coeffs for i in sys_.interfaces]
    inn = [i.int_coeffs for i in sys_.interfaces]
    d_source = torch.from_numpy(sys_.source).cuda()
    d_psi = torch.zeros(n_local, dtype=torch.float64, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        d_psi.zero_()
        torch.cuda.synchronize()
        return mat.solve_dev(ctl, d_psi.data_ptr(), d_source.data_ptr())

    def step_e2e():
        h_psi[:] = 0.0
        mat.set(h_diag, h_upper, None, bou, inn)
        import ctypes as C

        perf = capi.Perf()
        capi._check(capi.lib().b200ls_solve(mat.h, C.byref(ctl), C.c_void_p(h_psi.ctypes.data),
                                            C.c_void_p(h_source.ctypes.data), C.byref(perf)))
        return perf

    # ---- resident (value) ----
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, iters = 0.0, 0, 0
    for _ in range(args.steps):
        perf = step_resident()
        dev_ms += perf.solveMs + perf.setupMs
        launches += perf.kernelLaunches
        iters += perf.nIterations
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end (host buffers through the C-ABI) ----
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    wall_e2e = time.perf_counter() - t0

    # ---- kernel-level numbers for the roofline (CUDA events on the launching stream, inside the library) ----
    t_pre_ms = mat.time_kernel(1, 20)      # one DIC precondition = k_sweep_fwd + k_sweep_bwd
    t_amul_ms = mat.time_kernel(0, 50)

    times = torch.tensor([wall, wall_e2e, dev_ms / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall, wall_e2e, dev_s = [float(x) for x in times.cpu()]

    if rank == 0:
        peak, peak_src = load_peaks()
        its_per_step = iters / args.steps
        value = n_total * its_per_step * args.steps / wall
        e2e_value = n_total * its_per_step * args.steps / wall_e2e
        # algorithmic bytes: DIC precondition 72C + 32F ; Amul 24C + 16F (symmetric)
        b_pre = 72.0 * n_local + 32.0 * n_faces
        b_amul = 24.0 * n_local + 16.0 * n_faces
        ach = b_pre / (t_pre_ms * 1e-3) / 1e9
        # DRAM bytes of one fwd+bwd pair from the ncu --set full capture of this workload
        # (profiles/r01_ncu_full_pcg128_top_kernels.txt: 151.1 MB + 178.2 MB); only valid for the 128^3 subdomain
        traffic = 329.3e6 if n_local == N_SIDE ** 3 else None
        roofline = {"bound": "hbm", "kernel": "DIC precondition = k_sweep_fwd + k_sweep_bwd (wavefront sweeps)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "bytes_per_launch": b_pre, "ms_per_launch": t_pre_ms}
        spmv = {"kernel": "k_spmv (lduMatrix::Amul)", "achieved": b_amul / (t_amul_ms * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": b_amul / (t_amul_ms * 1e-3) / 1e9 / peak, "bytes_per_launch": b_amul,
                "ms_per_launch": t_amul_ms}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = run_reference_sample(sys_, 100)
                if r:
                    cpu = {"value": r[0], "unit": UNIT, "cores": 1, "kind": "reference",
                           "sample": f"{r[2]} PCG+DIC iterations on the same {N_SIDE}^3 matrix by the unmodified "
                                     f"reference solver (oracle/_ref, serial Pstream/dummy, {r[1]:.1f} s)"}
                    multi = run_reference_all_cores(N_SIDE, N_SIDE, N_SIDE, ITERS)
                    if multi and multi[0] > r[0]:
                        cpu = {"value": multi[0], "unit": UNIT, "cores": multi[1], "kind": "reference",
                               "sample": multi[2], "single_core_value": r[0]}
                else:
                    r = run_oracle_port_sample(sys_, 100)
                    cpu = {"value": r[0], "unit": UNIT, "cores": 1, "kind": "port",
                           "sample": f"{r[2]} PCG+DIC iterations on the same {N_SIDE}^3 matrix by the C restatement "
                                     f"oracle/ldu_oracle.c (oracle/_ref not built on this box, {r[1]:.1f} s)"}
            except Exception as ex:  # the reference binary may be absent on a box that never built it
                cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {ex}"}
        h2d = 8 * (n_local + n_faces) + 16 * n_local
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "n_cells": n_total, "iterations_per_step": its_per_step,
                       "l2": "working set 0.35 GB per GPU > 126 MB L2 (no flush needed)"},
            "ms_per_iteration": 1e3 * wall / args.steps / its_per_step,
            "device_ms_per_step": 1e3 * dev_s / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * n_local,
                    "ms_per_step": 1e3 * wall_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline, "spmv": spmv, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
