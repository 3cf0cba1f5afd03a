#!/usr/bin/env python3
"""Build the UNMODIFIED reference application icoFoam (test infrastructure only, OPTIONAL: not part of build()).

    python oracle/build_app.py [--ref /root/reference] [--jobs N]

Compiles, with plain g++ and ninja (no wmake: it needs flex), the reference's own sources where they lie:
libfileFormats, libsurfMesh, libtriSurface, libmeshTools, libfiniteVolume and
applications/legacy/incompressible/icoFoam (Make/options: finiteVolume + meshTools), on top of oracle/_ref/libOpenFOAM.so
(oracle/build_ref.py).  Outputs go to oracle/_app/ (git-ignored).  The two flex lexers (ASCII STL readers) are replaced by
oracle/stubs/stlLexerStubs.C.  icoFoam is the SURVEY 8(c) "full-application oracle": `pEqn.solve()` at icoFoam.C:116
with the shipped cavity case (PCG+DIC for p).  tests/golden/make_icofoam_golden.py drives it.
"""
import argparse
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from build_ref import CXXFLAGS, expand_make_files  # noqa: E402

OUT = HERE / "_app"
REFLIB = HERE / "_ref"
LIBS = ["fileFormats", "surfMesh", "triSurface", "meshTools", "finiteVolume"]
DEPS = {"fileFormats": [], "surfMesh": ["fileFormats"], "triSurface": ["fileFormats", "surfMesh"],
        "meshTools": ["triSurface", "surfMesh", "fileFormats"], "finiteVolume": ["triSurface", "meshTools", "surfMesh", "fileFormats"]}


def link_farm(src_root: Path, inc: Path):
    inc.mkdir(parents=True, exist_ok=True)
    for root, _dirs, files in os.walk(src_root):
        if "/lnInclude" in root or "/Make" in root:
            continue
        for f in files:
            if f.endswith((".H", ".C", ".h", ".T")):
                link = inc / f
                if not link.is_symlink():
                    try:
                        link.symlink_to(Path(root) / f)
                    except FileExistsError:
                        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 8)
    args = ap.parse_args()
    ref = Path(args.ref)
    if not (REFLIB / "libOpenFOAM.so").exists():
        print("oracle/_ref/libOpenFOAM.so missing: run python oracle/build_ref.py first", file=sys.stderr)
        return 1
    OUT.mkdir(exist_ok=True)
    (OUT / "obj").mkdir(exist_ok=True)
    lines = [
        f"cxxflags = {CXXFLAGS}",
        "rule cxx",
        "  command = g++ $cxxflags $inc -MMD -MF $out.d -c $in -o $out",
        "  depfile = $out.d",
        "  deps = gcc",
        "  description = CXX $out",
        "rule link",
        "  command = g++ -shared -o $out @$out.rsp $libs -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../_ref'",
        "  rspfile = $out.rsp",
        "  rspfile_content = $in",
        "  description = LINK $out",
        "rule exe",
        "  command = g++ -o $out $in $libs -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../_ref' -ldl -lm",
        "  description = EXE $out",
    ]
    n_units = 0
    for lib in LIBS:
        link_farm(ref / "src" / lib, OUT / f"lnInclude_{lib}")
        incs = " ".join(f"-I{OUT / ('lnInclude_' + d)}" for d in [lib] + DEPS[lib]) + f" -I{REFLIB / 'lnInclude'}"
        objs, seen = [], set()
        for rel in expand_make_files(ref / "src" / lib / "Make/files"):
            if rel.startswith("LIB"):
                continue
            if rel.endswith(".L"):
                continue          # flex source: stubbed below
            src = ref / "src" / lib / rel
            name = f"{lib}_{src.name.replace('.C', '')}"
            o = f"obj/{name}.o"
            k = 1
            while o in seen:
                o = f"obj/{name}_{k}.o"
                k += 1
            seen.add(o)
            objs.append(o)
            lines += [f"build {o}: cxx {src}", f"  inc = {incs}"]
            n_units += 1
        if lib == "triSurface":
            # both stubs live in one file; it needs surfMesh's and triSurface's headers
            o = "obj/stlLexerStubs.o"
            lines += [f"build {o}: cxx {HERE / 'stubs/stlLexerStubs.C'}", f"  inc = {incs}"]
            objs.append(o)
        deps = " ".join(f"lib{d}.so" for d in DEPS[lib] if not (lib == "surfMesh"))
        libs = f"-L{OUT} " + " ".join(f"-l{d}" for d in DEPS[lib]) + f" -L{REFLIB} -lOpenFOAM"
        if lib == "surfMesh":
            # its lexer stub is linked into libtriSurface; allow the undefined symbol here
            libs = f"-L{OUT} -lfileFormats -L{REFLIB} -lOpenFOAM"
            deps = "libfileFormats.so"
        lines += [f"build lib{lib}.so: link {' '.join(objs)} | {deps}", f"  libs = {libs}"]
    app = ref / "applications/legacy/incompressible/icoFoam"
    incs = f"-I{app} -I{OUT / 'lnInclude_finiteVolume'} -I{OUT / 'lnInclude_meshTools'} -I{REFLIB / 'lnInclude'}"
    lines += [f"build obj/icoFoam.o: cxx {app / 'icoFoam.C'}", f"  inc = {incs}",
              "build icoFoam: exe obj/icoFoam.o | " + " ".join(f"lib{d}.so" for d in LIBS),
              f"  libs = -L{OUT} -lfiniteVolume -lmeshTools -ltriSurface -lsurfMesh -lfileFormats -L{REFLIB} -lOpenFOAM",
              "default icoFoam"]
    (OUT / "build.ninja").write_text("\n".join(lines) + "\n")
    print(f"{n_units} compile units")
    rc = subprocess.run(["ninja", "-C", str(OUT), f"-j{args.jobs}", "-k", "20"]).returncode
    if rc == 0:
        # the libraries travel to the GPU box with the snapshot: drop the symbol tables they do not need
        subprocess.run(["strip", "--strip-unneeded"] + [str(OUT / f"lib{d}.so") for d in LIBS] + [str(OUT / "icoFoam")])
    return rc


if __name__ == "__main__":
    sys.exit(main())
