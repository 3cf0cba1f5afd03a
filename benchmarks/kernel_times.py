#!/usr/bin/env python3
"""Kernel-level timings (CUDA events inside the library, b200ls_time_kernel) and their algorithmic rooflines.

    python benchmarks/kernel_times.py 128 [256 ...]     -> one JSON line per size
Algorithmic bytes per SURVEY.md 8(d): Amul 24C+16F, DIC precondition 72C+32F, GaussSeidel sweep 60C+12F (symmetric).
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from _pkg import load_pkg  # noqa: E402

load_pkg()
from b200ls import capi, cases  # noqa: E402


def main():
    capi.init(0)
    peak = 6558.7
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    for n in [int(a) for a in sys.argv[1:]] or [128]:
        s = cases.cavity_laplacian(n, n, n)
        mesh, mat = capi.from_system(s)
        c, f = s.n_cells, s.n_faces
        rows = {}
        for name, which, nbytes, reps in (("Amul", 0, 24 * c + 16 * f, 50), ("DIC precondition (fwd+bwd sweeps)", 1, 72 * c + 32 * f, 20),
                                           ("GaussSeidel sweep (+ sentinel fill)", 2, 60 * c + 12 * f, 20)):
            ms = mat.time_kernel(which, reps)
            gbs = nbytes / (ms * 1e-3) / 1e9
            rows[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GB_per_s": gbs, "frac_of_measured_peak": gbs / peak}
        print(json.dumps({"n": n, "n_cells": c, "n_faces": f, "wavefronts": 3 * n - 2, "peak_GB_per_s": peak, "kernels": rows}))
        mat.close()
        mesh.close()


if __name__ == "__main__":
    main()
