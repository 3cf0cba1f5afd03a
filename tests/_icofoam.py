"""Helpers of the full-application parity test: run the reference's icoFoam (oracle/_app/icoFoam, built by
oracle/build_app.py from the unmodified sources) on a generated cavity case and parse its solver log."""
import os
import re
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ICOFOAM = ROOT / "oracle/_app/icoFoam"
PLUGIN = ROOT / "openfoam-dev_b200/libB200LinearSolvers.so"
_LINE = re.compile(r"^(\w+):\s+Solving for (\w+), Initial residual = (\S+), Final residual = (\S+), No Iterations (\d+)")


def run_icofoam(case_dir, timeout=600):
    env = dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM", WM_PROJECT_VERSION="dev")
    # run IN the case directory: its system/configDict (SolverPerformance 1 = print every solve) is merged over
    # etc/configDict only then (global/debug/debug.C:181-186)
    r = subprocess.run([str(ICOFOAM)], cwd=str(case_dir), env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(r.stdout[-3000:] + r.stderr[-3000:])
    return r.stdout


def parse_log(text):
    """[(solverName, field, initialResidual, finalResidual, nIterations), ...] in the order of the log."""
    out = []
    for line in text.splitlines():
        m = _LINE.match(line.strip())
        if m:
            out.append((m.group(1), m.group(2), float(m.group(3).rstrip(",")), float(m.group(4).rstrip(",")), int(m.group(5))))
    return out


def read_internal_field(path):
    """internalField of an ascii vol*Field file as an array [n] or [n, 3]."""
    text = Path(path).read_text()
    i = text.index("internalField")
    head = text[i: i + 200]
    if "nonuniform" not in head:
        raise ValueError(f"{path}: uniform internalField")
    j = text.index("(", i)
    n = int(text[i:j].split()[-1])
    k = text.index("\n)\n", j)
    body = text[j + 1: k].replace("(", " ").replace(")", " ")
    a = np.array(body.split(), dtype=np.float64)
    return a.reshape(n, -1).squeeze()
