#!/usr/bin/env python3
"""Build the UNMODIFIED reference solver stack into oracle/_ref/ (test infrastructure only).

This compiles the reference's own sources where they lie under /root/reference
(libOpenFOAM + OSspecific/POSIX + the serial Pstream/dummy backend) with plain g++,
without wmake (wmake needs flex, which this image lacks).  Nothing from the reference
is copied into the repository: the flat include directory is a directory of symlinks
and every output lands in oracle/_ref/ (git-ignored, NOT gpurun-ignored so the built
libOpenFOAM.so and harness travel to the GPU box).

Flags follow wmake/rules/linux64Gcc/c++:9-16 and wmake/rules/General/general:8-10.

Usage:  python oracle/build_ref.py [--ref /root/reference] [--jobs N]
"""
import argparse
import os
import re
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"

CXXFLAGS = (
    "-std=c++14 -m64 -Dlinux64 -DWM_ARCH_OPTION=64 -DWM_DP -DWM_LABEL_SIZE=32 "
    "-O3 -DNoRepository -ftemplate-depth-256 -fPIC -w"
)


def expand_make_files(path: Path):
    """Expand a wmake Make/files list: 'var = value' definitions and $(var) uses."""
    variables = {}
    out = []
    skip_depth = 0
    for raw in path.read_text().splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.startswith("ifeq") or line.startswith("ifneq"):
            # Only two conditionals occur: SP precision (we are DP) and SunOS64 (we are linux64).
            cond_true = False
            if "WM_PRECISION_OPTION" in line:
                cond_true = line.startswith("ifneq")  # DP != SP
            skip_depth = 0 if cond_true else 1
            continue
        if line == "else":
            skip_depth = 0 if skip_depth else 1
            continue
        if line == "endif":
            skip_depth = 0
            continue
        if skip_depth:
            continue
        m = re.match(r"^(\w+)\s*=\s*(.*)$", line)
        if m:
            val = m.group(2)
            val = re.sub(r"\$\((\w+)\)", lambda k: variables.get(k.group(1), k.group(0)), val)
            variables[m.group(1)] = val
            continue
        line = re.sub(r"\$\((\w+)\)", lambda k: variables.get(k.group(1), k.group(0)), line)
        out.append(line)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 8)
    args = ap.parse_args()
    ref = Path(args.ref)
    if not (ref / "src/OpenFOAM/Make/files").exists():
        print(f"reference not found at {ref}; nothing built", file=sys.stderr)
        return 1

    OUT.mkdir(exist_ok=True)
    inc = OUT / "lnInclude"
    obj = OUT / "obj"
    inc.mkdir(exist_ok=True)
    obj.mkdir(exist_ok=True)

    # 1. flat include dir (what wmakeLnInclude does) - symlinks only
    n = 0
    for top in (ref / "src/OpenFOAM", ref / "src/OSspecific/POSIX"):
        for root, _dirs, files in os.walk(top):
            if "/lnInclude" in root or "/Make" in root:
                continue
            for f in files:
                if f.endswith((".H", ".C", ".h", ".T")):
                    link = inc / f
                    if not link.is_symlink():
                        try:
                            link.symlink_to(Path(root) / f)
                            n += 1
                        except FileExistsError:
                            pass
    print(f"lnInclude: {n} new links")

    # 2. file list
    units = []  # (source path, extra flags)
    for rel in expand_make_files(ref / "src/OpenFOAM/Make/files"):
        if rel.endswith(".Cver"):
            gen = OUT / "global.Cver.C"
            text = (ref / "src/OpenFOAM" / rel).read_text()
            text = text.replace("VERSION_STRING", "dev").replace("BUILD_STRING", "oracle")
            if not gen.exists() or gen.read_text() != text:
                gen.write_text(text)
            units.append((gen, f"-I{ref / 'src/OpenFOAM/global'}"))
        else:
            units.append((ref / "src/OpenFOAM" / rel, ""))
    for rel in expand_make_files(ref / "src/OSspecific/POSIX/Make/files"):
        if rel == "dummyPrintStack.C":
            continue
        units.append((ref / "src/OSspecific/POSIX" / rel, "-DFOAM_USE_INOTIFY"))
    for f in ("UPstream.C", "UIPread.C", "UOPwrite.C"):
        units.append((ref / "src/Pstream/dummy" / f, ""))

    # 3. ninja build file
    lines = [
        f"cxxflags = {CXXFLAGS} -I{inc}",
        "rule cxx",
        "  command = g++ $cxxflags $extra -MMD -MF $out.d -c $in -o $out",
        "  depfile = $out.d",
        "  deps = gcc",
        "  description = CXX $out",
        "rule link",
        "  command = g++ -shared -o $out @$out.rsp -lz -ldl",
        "  rspfile = $out.rsp",
        "  rspfile_content = $in",
        "  description = LINK $out",
    ]
    objs = []
    seen = set()
    for src, extra in units:
        name = src.name.replace(".C", "")
        o = f"obj/{name}.o"
        k = 1
        while o in seen:
            o = f"obj/{name}_{k}.o"
            k += 1
        seen.add(o)
        objs.append(o)
        lines.append(f"build {o}: cxx {src}")
        if extra:
            lines.append(f"  extra = {extra}")
    lines.append("build libOpenFOAM.so: link " + " ".join(objs))
    lines.append("default libOpenFOAM.so")
    (OUT / "build.ninja").write_text("\n".join(lines) + "\n")
    print(f"{len(units)} compile units")

    r = subprocess.run(["ninja", "-C", str(OUT), f"-j{args.jobs}"])
    return r.returncode


if __name__ == "__main__":
    sys.exit(main())
