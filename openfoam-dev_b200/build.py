"""Build recipes (in-tree, no JIT cache): libb200ls.so (CUDA, sm_100a), the C oracle and, when the
reference tree is present, oracle/_ref (the unmodified reference solver stack + harness)."""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libb200ls.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                       # bit-parity with the reference's non-FMA x86-64 build
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-shared", "-cudart", "static",
]


def _newer(target: Path, sources):
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def build_lib(force=False, verbose=False):
    srcs = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.cpp")) +
                  list(CSRC.glob("*.hpp")) + [ROOT / "include" / "b200ls.h"])
    if not force and _newer(LIB, srcs):
        return LIB
    extra = os.environ.get("B200LS_EXTRA_NVCC_FLAGS", "").split()   # debugging builds (e.g. -DB200LS_PENCIL_FASTFAIL)
    cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [
        str(CSRC / "lib.cu"), str(CSRC / "mesh.cpp"), "-o", str(LIB), "-ldl",
    ]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def build_oracle(force=False):
    src = ROOT / "oracle" / "ldu_oracle.c"
    out = ROOT / "oracle" / "libldu_oracle.so"
    if not src.exists():
        return None
    if not force and _newer(out, [src]):
        return out
    # -ffp-contract=off: no FMA contraction, like the reference's x86-64 baseline build
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", str(src), "-o", str(out), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed on the oracle")
    return out


def build_ref(ref="/root/reference"):
    """oracle/_ref: only where the reference tree exists (the build container)."""
    if not Path(ref, "src/OpenFOAM/Make/files").exists():
        return None
    refdir = ROOT / "oracle" / "_ref"
    lib = refdir / "libOpenFOAM.so"
    if not lib.exists():
        subprocess.check_call([sys.executable, str(ROOT / "oracle" / "build_ref.py"), "--ref", ref])
    harness = refdir / "ref_harness"
    src = ROOT / "oracle" / "ref_harness.C"
    if not _newer(harness, [src, lib]):
        subprocess.check_call([
            "g++", "-std=c++14", "-m64", "-Dlinux64", "-DWM_ARCH_OPTION=64", "-DWM_DP", "-DWM_LABEL_SIZE=32",
            "-O3", "-DNoRepository", "-ftemplate-depth-256", "-w", f"-I{refdir / 'lnInclude'}",
            str(src), "-o", str(harness), f"-L{refdir}", "-lOpenFOAM", "-Wl,-rpath,$ORIGIN", "-ldl",
        ])
    return harness


def build_plugin(ref="/root/reference"):
    """libB200LinearSolvers.so: the OpenFOAM plugin shim, compiled against the reference headers (build container
    only; the built .so travels to the GPU box together with oracle/_ref/libOpenFOAM.so)."""
    refdir = ROOT / "oracle" / "_ref"
    if not (refdir / "lnInclude").exists() or not Path(ref, "src/OpenFOAM").exists():
        return None
    src = PKG / "plugin" / "B200Solvers.C"
    out = PKG / "libB200LinearSolvers.so"
    if _newer(out, [src, ROOT / "include" / "b200ls.h", LIB]):
        return out
    subprocess.check_call([
        "g++", "-std=c++14", "-m64", "-Dlinux64", "-DWM_ARCH_OPTION=64", "-DWM_DP", "-DWM_LABEL_SIZE=32", "-O2",
        "-DNoRepository", "-ftemplate-depth-256", "-fPIC", "-w", f"-I{refdir / 'lnInclude'}", f"-I{ROOT / 'include'}",
        "-shared", str(src), "-o", str(out), f"-L{PKG}", "-lb200ls", f"-L{refdir}", "-lOpenFOAM",
        "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../oracle/_ref",
    ])
    return out


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle()
    print(LIB)
