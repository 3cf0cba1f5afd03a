#!/usr/bin/env python3
"""Debugging aid for the pencil sweeps: run a few DIC preconditions on an n^3 block with B200LS_PENCIL_PROF set and
summarise the per-tile counters of the last forward / backward launch.

    python benchmarks/pencil_prof.py 128 [out.txt]
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prof = sys.argv[2] if len(sys.argv) > 2 else tempfile.mktemp()
os.environ["B200LS_PENCIL_PROF"] = prof
from _pkg import load_pkg  # noqa: E402

load_pkg()
from b200ls import capi, cases  # noqa: E402

capi.init(0)
s = cases.cavity_laplacian(n, n, n)
mesh, mat = capi.from_system(s)
which = int(os.environ.get("PROF_WHICH", "1"))
ms = mat.time_kernel(which, 3)
print("ms per call (with the profiling syncs):", ms)
blocks = open(prof).read().split("# ")[1:]
for blk in blocks[-2:]:
    head, *rows = blk.strip().split("\n")
    a = np.array([[int(v) for v in r.split()] for r in rows], dtype=np.int64)
    t0 = a[:, 0].min()
    start, end = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3
    print(head)
    print(f"  kernel span {end.max():.1f} us; tile duration mean {np.mean(end - start):.1f} us, min {np.min(end - start):.1f}, "
          f"max {np.max(end - start):.1f}")
    names = ["chain: wait record", "chain: wait result ring", "prep: wait operands", "prep: wait neighbour values",
             "prep: wait record slot", "writer: wait results", "helper: polling rounds", "helper: wait ring capacity"]
    for k, nm in enumerate(names):
        print(f"  {nm:28s} mean {a[:, 2 + k].mean():10.0f}  max {a[:, 2 + k].max():10d}")
    # time at which the chain warp starts each group of eight steps, for the first tiles
    for i in [0, 1, 2, 3, 4, 17]:
        g = a[i, 16:16 + 19]
        print(f"  tile {i:3d} group starts (us): " + " ".join(f"{(v - t0) / 1e3:6.1f}" if v else "   -  " for v in g))
    mhz = a[:, 10] / np.maximum(1e-9, (a[:, 1] - a[:, 0]) / 1e3)
    print(f"  chain warp: SM clock seen = cycles / globaltimer: median {np.median(mhz):.0f} MHz (min {mhz.min():.0f}, max {mhz.max():.0f})")
    for i in [0, 1, 2, 3, 16, 100, 255, 300, len(a) - 1]:
        if i < len(a):
            print(f"  tile {i:4d}: start {start[i]:8.1f} end {end[i]:8.1f} us | " + " ".join(f"{v:8d}" for v in a[i, 2:10]))
