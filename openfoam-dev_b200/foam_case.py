"""Writes a complete OpenFOAM case of the lid-driven cavity for the reference's icoFoam
(applications/legacy/incompressible/icoFoam): polyMesh of an nx*ny*nz hex block numbered as cases.block_addressing
numbers it (cells i-fastest, internal faces in upper-triangular order, the layout blockMesh produces for one block),
fields, schemes and solver controls with the values of the shipped tutorial
(tutorials/legacy/incompressible/icoFoam/cavity/cavity: nu 0.01, deltaT 0.005, lid speed 1, PISO with 2 correctors,
p: PCG+DIC 1e-06/0.05, U: smoothSolver symGaussSeidel 1e-05/0).  Used by the full-application parity test: the same case
runs once with the reference's solvers and once with `libs ("libB200LinearSolvers.so")` + the B200 solver names."""
from pathlib import Path

import numpy as np

_HDR = """FoamFile
{{
    format      ascii;
    class       {cls};
    location    "{loc}";
    object      {obj};
}}

"""


def _write(path, cls, body):
    path.parent.mkdir(parents=True, exist_ok=True)
    loc = path.parent.name if path.parent.name != "polyMesh" else "constant/polyMesh"
    path.write_text(_HDR.format(cls=cls, loc=loc, obj=path.name) + body)


def _list(items, fmt):
    return f"{len(items)}\n(\n" + "\n".join(fmt(x) for x in items) + "\n)\n"


def write_block_polymesh(mesh_dir, nx, ny, nz, lx=0.1, ly=0.1, lz=0.01, cyclic_z=False):
    """points / faces / owner / neighbour / boundary of the block; patches movingWall (y max), fixedWalls (x min,
    x max, y min), frontAndBack (z min, z max; `empty` when nz == 1) -- or, with cyclic_z, the cyclic pair
    front (z min) / back (z max), face i of one coupled to face i of the other."""
    mesh_dir = Path(mesh_dir)
    P = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)   # noqa: E731
    C = lambda i, j, k: i + nx * (j + ny * k)               # noqa: E731
    pts = [(lx * i / nx, ly * j / ny, lz * k / nz) for k in range(nz + 1) for j in range(ny + 1) for i in range(nx + 1)]
    faces, owner, neigh = [], [], []
    fx = lambda i, j, k: (P(i, j, k), P(i, j + 1, k), P(i, j + 1, k + 1), P(i, j, k + 1))      # +x normal  # noqa: E731
    fy = lambda i, j, k: (P(i, j, k), P(i, j, k + 1), P(i + 1, j, k + 1), P(i + 1, j, k))      # +y normal  # noqa: E731
    fz = lambda i, j, k: (P(i, j, k), P(i + 1, j, k), P(i + 1, j + 1, k), P(i, j + 1, k))      # +z normal  # noqa: E731
    rev = lambda f: (f[0], f[3], f[2], f[1])                                                    # noqa: E731
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = C(i, j, k)
                if i < nx - 1:
                    faces.append(fx(i + 1, j, k)); owner.append(c); neigh.append(c + 1)
                if j < ny - 1:
                    faces.append(fy(i, j + 1, k)); owner.append(c); neigh.append(c + nx)
                if k < nz - 1:
                    faces.append(fz(i, j, k + 1)); owner.append(c); neigh.append(c + nx * ny)
    n_int = len(faces)
    patches = []

    def patch(name, typ, items):
        start = len(faces)
        for f, c in items:
            faces.append(f); owner.append(c)
        patches.append((name, typ, start, len(items)))

    patch("movingWall", "wall", [(fy(i, ny, k), C(i, ny - 1, k)) for k in range(nz) for i in range(nx)])
    patch("fixedWalls", "wall",
          [(rev(fx(0, j, k)), C(0, j, k)) for k in range(nz) for j in range(ny)]
          + [(fx(nx, j, k), C(nx - 1, j, k)) for k in range(nz) for j in range(ny)]
          + [(rev(fy(i, 0, k)), C(i, 0, k)) for k in range(nz) for i in range(nx)])
    if cyclic_z:
        patch("front", "cyclic", [(rev(fz(i, j, 0)), C(i, j, 0)) for j in range(ny) for i in range(nx)])
        patch("back", "cyclic", [(fz(i, j, nz), C(i, j, nz - 1)) for j in range(ny) for i in range(nx)])
    else:
        patch("frontAndBack", "empty" if nz == 1 else "wall",
              [(rev(fz(i, j, 0)), C(i, j, 0)) for j in range(ny) for i in range(nx)]
              + [(fz(i, j, nz), C(i, j, nz - 1)) for j in range(ny) for i in range(nx)])
    _write(mesh_dir / "points", "vectorField", _list(pts, lambda p: f"({p[0]:.17g} {p[1]:.17g} {p[2]:.17g})"))
    _write(mesh_dir / "faces", "faceList", _list(faces, lambda f: f"4({f[0]} {f[1]} {f[2]} {f[3]})"))
    _write(mesh_dir / "owner", "labelList", _list(owner, str))
    _write(mesh_dir / "neighbour", "labelList", _list(neigh, str))
    body = f"{len(patches)}\n(\n"
    for name, typ, start, n in patches:
        body += f"    {name}\n    {{\n        type            {typ};\n"
        if typ == "wall":
            body += "        inGroups        List<word> 1(wall);\n"
        if typ == "cyclic":
            body += ("        inGroups        List<word> 1(cyclic);\n"
                     f"        neighbourPatch  {'back' if name == 'front' else 'front'};\n")
        body += f"        nFaces          {n};\n        startFace       {start};\n    }}\n"
    body += ")\n"
    _write(mesh_dir / "boundary", "polyBoundaryMesh", body)
    return n_int, np.array(owner[:n_int], dtype=np.int32), np.array(neigh, dtype=np.int32)


def write_cavity_case(case_dir, nx=20, ny=20, nz=1, p_solver="solver PCG; preconditioner DIC; tolerance 1e-06; relTol 0.05;",
                      p_final="$p; relTol 0;", u_solver="solver smoothSolver; smoother symGaussSeidel; tolerance 1e-05; relTol 0;",
                      libs=None, end_time=0.05, delta_t=0.005, write=False, cyclic_z=False, lz=0.01):
    case_dir = Path(case_dir)
    fb = "empty" if nz == 1 else "noSlip"
    write_block_polymesh(case_dir / "constant/polyMesh", nx, ny, nz, lz=lz, cyclic_z=cyclic_z)
    _write(case_dir / "constant/physicalProperties", "dictionary", "nu              [0 2 -1 0 0 0 0] 0.01;\n")
    _write(case_dir / "0/U", "volVectorField",
           "dimensions      [0 1 -1 0 0 0 0];\ninternalField   uniform (0 0 0);\nboundaryField\n{\n"
           "    movingWall { type fixedValue; value uniform (1 0 0); }\n    fixedWalls { type noSlip; }\n"
           + (f"    frontAndBack {{ type {fb}; }}\n}}\n" if not cyclic_z else
              "    front { type cyclic; }\n    back { type cyclic; }\n}\n"))
    fbp = "empty" if nz == 1 else "zeroGradient"
    _write(case_dir / "0/p", "volScalarField",
           "dimensions      [0 2 -2 0 0 0 0];\ninternalField   uniform 0;\nboundaryField\n{\n"
           "    movingWall { type zeroGradient; }\n    fixedWalls { type zeroGradient; }\n"
           + (f"    frontAndBack {{ type {fbp}; }}\n}}\n" if not cyclic_z else
              "    front { type cyclic; }\n    back { type cyclic; }\n}\n"))
    libs_line = f"libs ({libs});\n" if libs else ""
    _write(case_dir / "system/controlDict", "dictionary",
           f"application     icoFoam;\n{libs_line}startFrom       startTime;\nstartTime       0;\nstopAt          endTime;\n"
           f"endTime         {end_time};\ndeltaT          {delta_t};\nwriteControl    timeStep;\n"
           f"writeInterval   {int(round(end_time / delta_t)) if write else 100000000};\npurgeWrite      0;\nwriteFormat     ascii;\n"
           "writePrecision  16;\nwriteCompression off;\ntimeFormat      general;\ntimePrecision   6;\n"
           "runTimeModifiable false;\n")
    # merged over etc/configDict when the application runs in the case directory (global/debug/debug.C:181-186):
    # print every linear solve
    (case_dir / "system/configDict").write_text("DebugSwitches { SolverPerformance 1; }\n")
    _write(case_dir / "system/fvSchemes", "dictionary",
           "ddtSchemes { default Euler; }\ngradSchemes { default Gauss linear; }\n"
           "divSchemes { default none; div(phi,U) Gauss linear; }\nlaplacianSchemes { default Gauss linear orthogonal; }\n"
           "interpolationSchemes { default linear; }\nsnGradSchemes { default orthogonal; }\n")
    _write(case_dir / "system/fvSolution", "dictionary",
           f"solvers\n{{\n    p {{ {p_solver} }}\n    pFinal {{ {p_final} }}\n    U {{ {u_solver} }}\n}}\n"
           "PISO\n{\n    nCorrectors     2;\n    nNonOrthogonalCorrectors 0;\n    pRefCell        0;\n    pRefValue       0;\n}\n")
    return case_dir
