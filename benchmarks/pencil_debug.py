#!/usr/bin/env python3
"""Debugging aid: where does a pencil operator first go wrong?  Reports the mismatching cells per tile and, for the
tile that is wrong first in processing order, per (i, jj, kk)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle"))
os.environ.setdefault("B200LS_PENCIL_MIN_CELLS", "0")
from _pkg import load_pkg  # noqa: E402

load_pkg()
import ldu_oracle as orc  # noqa: E402
from b200ls import capi, cases  # noqa: E402

TRACE = os.environ.get("B200LS_PENCIL_TRACE")
if TRACE and os.path.exists(TRACE):
    os.remove(TRACE)
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "24,16,8").split(","))
op = sys.argv[2] if len(sys.argv) > 2 else "symGaussSeidel"
nsw = int(sys.argv[3]) if len(sys.argv) > 3 else 1
capi.init(0)
nx, ny, nz = shape
s = cases.cavity_laplacian(nx, ny, nz, coeffs="random")
mesh, mat = capi.from_system(s)
S = orc.System(s)
x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
got = mat.smooth(op, x, s.source, nsw)
want = orc.smooth(S, op, x, s.source, nsw)
bad = ~(got == want)
if TRACE:
    raw = open(TRACE, "rb").read()
    pos = 0
    while pos < len(raw):
        mode, n_tiles, tnx, _ = np.frombuffer(raw, np.int32, 4, pos)
        pos += 16
        t = np.frombuffer(raw, np.float64, 4 * n_tiles * 4096, pos).reshape(4, n_tiles, 4096)
        pos += 8 * 4 * n_tiles * 4096
        for ti in range(n_tiles):
            S = tnx + 10
            prep, chain, y, wr = t[0, ti, :S], t[1, ti, :S], t[2, ti, :S], t[3, ti, :S]
            d1 = np.flatnonzero(prep.view(np.int64) != chain.view(np.int64))
            d2 = np.flatnonzero(y.view(np.int64) != wr.view(np.int64))
            print(f"trace mode {mode} tile(launch order) {ti}: prep!=chain at steps {d1[:8]}, chain-result!=writer at steps {d2[:8]}")
            for k in d1[:3]:
                print(f"    step {k}: prep wrote {prep[k]!r}, chain took {chain[k]!r}; prep[k-8..k+8] matches chain value at "
                      f"{[int(j) for j in range(max(0, k - 8), min(S, k + 9)) if prep[j] == chain[k]]}")
print(shape, op, nsw, "bad cells:", int(bad.sum()), "nan:", int(np.isnan(got).sum()))
dims = mesh.get_i32(21, 0)
WJ, WK, nJ, nK = dims[3:7]
B = bad.reshape(nz, ny, nx)
G = got.reshape(nz, ny, nx)
for K in range(nK):
    print("K", K, " ".join(f"{int(B[K * WK:(K + 1) * WK, J * WJ:(J + 1) * WJ, :].sum()):6d}" for J in range(nJ)))
# the first tile in forward (wavefront) order that has a bad cell, and its earliest bad cells
best = None
for K in range(nK):
    for J in range(nJ):
        blk = B[K * WK:(K + 1) * WK, J * WJ:(J + 1) * WJ, :]
        if blk.any() and (best is None or J + K < best[0]):
            best = (J + K, J, K)
if best:
    _, J, K = best
    blk = B[K * WK:(K + 1) * WK, J * WJ:(J + 1) * WJ, :]
    kk, jj, ii = np.nonzero(blk)
    order = np.argsort(ii + jj + kk)
    print("first bad tile in forward order: J", J, "K", K, "bad", int(blk.sum()))
    W3 = want.reshape(nz, ny, nx)
    for t in order[:10]:
        k, j, i = kk[t] + K * WK, jj[t] + J * WJ, ii[t]
        print(f"   step {ii[t] + jj[t] + kk[t]:3d} i {ii[t]} jj {jj[t]} kk {kk[t]}: got {G[k, j, i]!r} want {W3[k, j, i]!r}")
# the last tile in forward order that has a bad cell = the first one in reverse order
for K in reversed(range(nK)):
    for J in reversed(range(nJ)):
        blk = B[K * WK:(K + 1) * WK, J * WJ:(J + 1) * WJ, :]
        if blk.any():
            print("first bad tile in reverse order: J", J, "K", K)
            kk, jj, ii = np.nonzero(blk)
            order = np.argsort(-(ii * 10000 + jj * 100 + kk))
            for t in order[:12]:
                k, j, i = kk[t] + K * WK, jj[t] + J * WJ, ii[t]
                print(f"   i {ii[t]} jj {jj[t]} kk {kk[t]}: got {G[k, j, i]!r} want {want.reshape(nz, ny, nx)[k, j, i]!r}")
            raise SystemExit
