"""CPU tests (no GPU): host-side integer data of libb200ls against the reference goldens (bit-exact), the C-ABI
surface, the B2LS container and the decomposition helpers."""
import re
from pathlib import Path

import numpy as np
import pytest

from _util import FIXTURES, min_cells_of, capi, cases, ldu_io, load_fixture, system_from_entries

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "b200ls.h").read_text()
    declared = set(re.findall(r"\b(b200ls_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200ls_mesh_s", "b200ls_matrix_s"}
    lib = capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200ls.h but not exported"
    assert set(capi.EXPORTS) == declared


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when no CUDA device is usable."""
    if capi.device_available():
        pytest.skip("a CUDA device is present")
    s = cases.cavity_laplacian(4, 4, 1)
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper)
    mat = capi.Matrix(mesh)
    with pytest.raises(capi.B200Error, match="no CUDA device"):
        mat.set(s.diag, s.upper_coeffs)


@pytest.mark.parametrize("name", FIXTURES)
def test_addressing_bit_exact(name):
    inp, ref = load_fixture(name)
    s = system_from_entries(inp)
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper)
    assert np.array_equal(mesh.get_i32(capi.LOSORT), ref["losort"])
    assert np.array_equal(mesh.get_i32(capi.OWNER_START), ref["ownerStart"])
    assert np.array_equal(mesh.get_i32(capi.LOSORT_START), ref["losortStart"])


@pytest.mark.parametrize("name", [f for f in FIXTURES if f != "block_24x24x24"])
def test_agglomeration_bit_exact(name):
    inp, ref = load_fixture(name)
    s = system_from_entries(inp)
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper, s.interfaces)
    n_coarse = mesh.agglomerate(s.face_weights, min_cells_per_processor=min_cells_of(inp), forward_start=1)
    assert n_coarse == int(ref["agg.nLevels"][0])
    assert mesh.n_levels == n_coarse + 1
    for lev in range(n_coarse):
        k = f"agg.{lev}."
        # coarse cyclic patches (cyclicGAMGInterface.C:85-157), from the unmodified reference
        for i in range(len(s.interfaces)):
            assert np.array_equal(mesh.get_iface_i32(0, lev + 1, i), ref[f"{k}iface.{i}.faceCells"]), (lev, i)
            assert np.array_equal(mesh.get_iface_i32(1, lev, i), ref[f"{k}iface.{i}.faceRestrictAddressing"]), (lev, i)
        assert np.array_equal(mesh.get_i32(capi.RESTRICT_ADDRESSING, lev), ref[k + "restrictAddressing"])
        assert np.array_equal(mesh.get_i32(capi.FACE_RESTRICT_ADDRESSING, lev), ref[k + "faceRestrictAddressing"])
        assert np.array_equal(mesh.get_i32(capi.FACE_FLIP_MAP, lev), ref[k + "faceFlipMap"].astype(np.int32))
        assert np.array_equal(mesh.get_i32(capi.LOWER_ADDR, lev + 1), ref[k + "coarseLower"])
        assert np.array_equal(mesh.get_i32(capi.UPPER_ADDR, lev + 1), ref[k + "coarseUpper"])
        assert list(mesh.get_i32(capi.LEVEL_SIZES, lev + 1)) == list(ref[k + "sizes"])


@pytest.mark.parametrize("shape", [(5, 4, 3), (20, 20, 1), (9, 1, 1), (1, 1, 1)])
def test_wavefronts_of_a_block(shape):
    """Canonical wavefronts: on an nx*ny*nz block Lf = i+j+k and Lb mirrors it."""
    nx, ny, nz = shape
    lower, upper, _ = cases.block_addressing(nx, ny, nz)
    n = nx * ny * nz
    mesh = capi.Mesh(n, lower, upper)
    c = np.arange(n)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    for which_o, which_r, lev in ((capi.FWD_LEVEL_OFFSETS, capi.FWD_LEVEL_ROWS, i + j + k),
                                  (capi.BWD_LEVEL_OFFSETS, capi.BWD_LEVEL_ROWS,
                                   (nx - 1 - i) + (ny - 1 - j) + (nz - 1 - k))):
        offs, rows = mesh.get_i32(which_o), mesh.get_i32(which_r)
        assert offs.size - 1 == nx + ny + nz - 2
        for q in range(offs.size - 1):
            assert np.array_equal(rows[offs[q]:offs[q + 1]], c[lev == q])


def test_wavefront_dependencies_general_mesh():
    """Every row sits strictly after all of its dependencies, on an unstructured (agglomerated) level."""
    s = cases.cavity_laplacian(11, 7, 5, coeffs="random")
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper)
    mesh.agglomerate(s.face_weights)
    for lev in range(mesh.n_levels):
        lo, up = mesh.get_i32(capi.LOWER_ADDR, lev), mesh.get_i32(capi.UPPER_ADDR, lev)
        n = int(mesh.get_i32(capi.LEVEL_SIZES, lev)[0])
        for wo, wr, src, dst in ((capi.FWD_LEVEL_OFFSETS, capi.FWD_LEVEL_ROWS, lo, up),
                                 (capi.BWD_LEVEL_OFFSETS, capi.BWD_LEVEL_ROWS, up, lo)):
            offs, rows = mesh.get_i32(wo, lev), mesh.get_i32(wr, lev)
            level_of = np.empty(n, dtype=np.int64)
            for q in range(offs.size - 1):
                level_of[rows[offs[q]:offs[q + 1]]] = q
            assert sorted(rows.tolist()) == list(range(n))
            assert np.all(level_of[dst] > level_of[src])
            # tight: every row of level q>0 has a dependency in level q-1
            need = np.zeros(n, dtype=np.int64)
            np.maximum.at(need, dst, level_of[src] + 1)
            assert np.array_equal(need, level_of)


def test_invalid_addressing_is_rejected():
    with pytest.raises(capi.B200Error, match="upper-triangular"):
        capi.Mesh(4, [1, 0], [2, 3])
    with pytest.raises(capi.B200Error, match="lower < upper"):
        capi.Mesh(4, [2], [1])


def test_b2ls_roundtrip(tmp_path):
    e = {"a": np.arange(5, dtype=np.int32), "b": np.linspace(0, 1, 7), "s": "solver PCG;", "n": 3}
    ldu_io.write(str(tmp_path / "x.b2ls"), e)
    r = ldu_io.read(str(tmp_path / "x.b2ls"))
    assert np.array_equal(r["a"], e["a"]) and np.array_equal(r["b"], e["b"])
    assert ldu_io.as_str(r["s"]) == "solver PCG;" and int(r["n"][0]) == 3


@pytest.mark.parametrize("n_ranks", [2, 4, 8])
def test_direct_subdomain_equals_general_decomposition(n_ranks):
    from b200ls import decompose

    split = decompose.simple_split(n_ranks)
    nx, ny, nz = 4 * split[0], 3 * split[1], 2 * split[2]
    glob = cases.cavity_laplacian(nx, ny, nz)
    parts, maps = decompose.decompose_system(glob, decompose.box_cell_ranks(nx, ny, nz, split), n_ranks)
    for r in range(n_ranks):
        d = decompose.cavity_subdomain(nx, ny, nz, split, r)
        g = parts[r]
        assert d.n_cells == g.n_cells
        assert np.array_equal(d.lower, g.lower) and np.array_equal(d.upper, g.upper)
        assert np.array_equal(d.diag, g.diag) and np.array_equal(d.upper_coeffs, g.upper_coeffs)
        assert np.allclose(d.source, g.source, rtol=0, atol=0)
        assert [i.neighb_rank for i in d.interfaces] == [i.neighb_rank for i in g.interfaces]
        for a, b in zip(d.interfaces, g.interfaces):
            assert np.array_equal(a.face_cells, b.face_cells) and np.array_equal(a.bou_coeffs, b.bou_coeffs)


def test_decomposed_operator_equals_global_operator():
    """numpy restatement of Amul with interfaces (lduMatrixATmul.C:34-92 + processorFvPatchScalarField.C:133-136)
    on the decomposed systems reproduces the global Amul."""
    from b200ls import decompose

    glob = cases.convection_diffusion(6, 4, 4)
    ranks = decompose.box_cell_ranks(6, 4, 4, (2, 2, 1))
    parts, maps = decompose.decompose_system(glob, ranks, 4)
    x = np.cos(0.3 * np.arange(glob.n_cells))

    def amul(s, xl):
        y = s.diag * xl
        lo_c = s.upper_coeffs if s.lower_coeffs is None else s.lower_coeffs
        np.add.at(y, s.upper, lo_c * xl[s.lower])
        np.add.at(y, s.lower, s.upper_coeffs * xl[s.upper])
        return y

    yg = amul(glob, x)
    for r, (s, m) in enumerate(zip(parts, maps)):
        y = amul(s, x[m])
        for itf in s.interfaces:
            # the neighbour's patchInternalField in the same face order
            other = parts[itf.neighb_rank]
            back = [i for i in other.interfaces if i.neighb_rank == r][0]
            recv = x[maps[itf.neighb_rank]][back.face_cells]
            np.subtract.at(y, itf.face_cells, itf.bou_coeffs * recv)
        assert np.allclose(y, yg[m], rtol=1e-14, atol=1e-14)


def test_agglomerate_from_reference_maps():
    """Levels rebuilt from the reference's restrictAddressing alone reproduce its face maps and coarse addressing."""
    inp, ref = load_fixture("block_16x16x16_rand")
    s = system_from_entries(inp)
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper)
    n = int(ref["agg.nLevels"][0])
    assert mesh.agglomerate_from_maps([ref[f"agg.{k}.restrictAddressing"] for k in range(n)]) == n
    for lev in range(n):
        k = f"agg.{lev}."
        assert np.array_equal(mesh.get_i32(capi.FACE_RESTRICT_ADDRESSING, lev), ref[k + "faceRestrictAddressing"])
        assert np.array_equal(mesh.get_i32(capi.FACE_FLIP_MAP, lev), ref[k + "faceFlipMap"].astype(np.int32))
        assert np.array_equal(mesh.get_i32(capi.LOWER_ADDR, lev + 1), ref[k + "coarseLower"])
        assert np.array_equal(mesh.get_i32(capi.UPPER_ADDR, lev + 1), ref[k + "coarseUpper"])


@pytest.mark.parametrize("seed", range(6))
def test_agglomeration_matches_c_oracle_on_random_graphs(seed):
    """Two independent restatements (C++ in libb200ls, plain C in oracle/) of the pair agglomeration and of
    agglomerateLduAddressing agree level by level on irregular graphs (both are pinned to the reference on fixtures)."""
    import sys

    sys.path.insert(0, str(ROOT / "oracle"))
    import ldu_oracle as orc

    rng = np.random.default_rng(seed)
    s = cases.random_graph(int(rng.integers(60, 900)), avg_degree=int(rng.integers(2, 9)), symmetric=bool(seed % 2),
                           seed=100 + seed, max_span=int(rng.integers(5, 200)))
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper)
    n = mesh.agglomerate(s.face_weights, forward_start=1)
    levels = orc.agglomeration(orc.System(s))
    assert n == len(levels)
    for k, (ra, fra, ff, cl, cu) in enumerate(levels):
        assert np.array_equal(mesh.get_i32(capi.RESTRICT_ADDRESSING, k), ra)
        assert np.array_equal(mesh.get_i32(capi.FACE_RESTRICT_ADDRESSING, k), fra)
        assert np.array_equal(mesh.get_i32(capi.FACE_FLIP_MAP, k), ff)
        assert np.array_equal(mesh.get_i32(capi.LOWER_ADDR, k + 1), cl)
        assert np.array_equal(mesh.get_i32(capi.UPPER_ADDR, k + 1), cu)


def test_forward_flag_alternates_like_the_reference_static():
    """pairGAMGAgglomeration::forward_ is process-global in the reference: with forwardStart = -1 the library
    mirrors that (a second mesh starts with the flag the first one left), with 0/1 it is explicit."""
    s = cases.random_graph(300, symmetric=True, seed=11)     # irregular: visiting order matters
    a = capi.Mesh(s.n_cells, s.lower, s.upper)
    b = capi.Mesh(s.n_cells, s.lower, s.upper)
    a.agglomerate(s.face_weights, forward_start=1)
    b.agglomerate(s.face_weights, forward_start=0)
    assert not np.array_equal(a.get_i32(capi.RESTRICT_ADDRESSING, 0), b.get_i32(capi.RESTRICT_ADDRESSING, 0))
    c = capi.Mesh(s.n_cells, s.lower, s.upper)
    c.agglomerate(s.face_weights, forward_start=1)
    assert np.array_equal(a.get_i32(capi.RESTRICT_ADDRESSING, 0), c.get_i32(capi.RESTRICT_ADDRESSING, 0))


def test_cyclic_patch_validation():
    lower, upper, _ = cases.block_addressing(4, 3, 2)
    a = cases.Interface(neighb_rank=-1, face_cells=np.array([0, 4, 8], np.int32), bou_coeffs=np.ones(3),
                        int_coeffs=np.ones(3), nbr_patch=1)
    b = cases.Interface(neighb_rank=-1, face_cells=np.array([3, 7], np.int32), bou_coeffs=np.ones(2),
                        int_coeffs=np.ones(2), nbr_patch=0)
    with pytest.raises(capi.B200Error, match="different size"):
        capi.Mesh(24, lower, upper, [a, b])
    b.face_cells = np.array([3, 7, 11], np.int32)
    b.nbr_patch = 1
    with pytest.raises(capi.B200Error, match="point back"):
        capi.Mesh(24, lower, upper, [a, b])
    b.nbr_patch = 0
    capi.Mesh(24, lower, upper, [a, b])


def _emulate_pencil_precondition(mesh, s, rA, skew=2):
    """Executes the pencil schedule (mesh.hpp PencilPlan, csrc/pencil.cuh) step by step: tiles in launch order, lane =
    pencil, lane (jj, kk) skewed by skew*(jj + kk) steps; the lower neighbours of a row are the lane's own previous
    result, the results two neighbouring lanes produced `skew` steps earlier, or -- on the low faces of a tile --
    values of tiles (J-1, K) / (J, K-1), which must already be complete (asserted).  The backward sweep runs the
    same schedule on reflected coordinates.  Arithmetic order as in k_pencil.  Returns wA in cell order."""
    g = lambda w: mesh.get_i32(w, 0)   # noqa: E731
    perm = g(13)
    nx, ny, nz, WJ, WK, nJ, nK = g(21)
    tiles = g(22).reshape(-1, 10)
    order = g(23)
    lptr, lface, uptr, uface = g(14), g(16), g(17), g(19)
    n = s.n_cells
    import sys as _sys
    _sys.path.insert(0, str(ROOT / "oracle"))
    import ldu_oracle as orc
    rD = orc.reciprocal_d(orc.System(s))[perm]
    lower_c = s.upper_coeffs if s.lower_coeffs is None else s.lower_coeffs
    # coefficient planes exactly as k_pencil_planes / k_pencil_pack build them
    tL = np.zeros((3, n))
    tU = np.zeros((3, n))
    ijk = np.empty((n, 3), np.int64)
    for p in range(n):
        c = perm[p]
        i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
        ijk[p] = (i, j, k)
        e = lptr[p]
        for slot, has in enumerate((k > 0, j > 0, i > 0)):
            if has:
                tL[slot, p] = rD[p] * lower_c[lface[e]]
                e += 1
        assert e == lptr[p + 1]
        e = uptr[p]
        cu = [0.0, 0.0, 0.0]
        for slot, has in enumerate((i < nx - 1, j < ny - 1, k < nz - 1)):
            if has:
                cu[slot] = s.upper_coeffs[uface[e]]
                e += 1
        assert e == uptr[p + 1]
        for slot, has in enumerate((k < nz - 1, j < ny - 1, i < nx - 1)):
            if has:
                tU[slot, p] = rD[p] * cu[2 - slot]
    fin = rA[perm]
    y = np.full(n, np.nan)
    z = np.full(n, np.nan)
    for direction, src, dst, t in ((1, fin, y, tL), (-1, y, z, tU)):
        done_tiles = set()
        for ti in (order if direction > 0 else order[::-1]):
            base, w, wj, wk, j0, k0, nJm, nKm, nJp, nKp = tiles[ti]
            nbJ, nbK = (nJm, nKm) if direction > 0 else (nJp, nKp)
            for nb in (nbJ, nbK):
                assert nb < 0 or nb in done_tiles, "neighbour tile not launched earlier"
            hist = {}   # step -> per-lane value
            S = nx + skew * (wj - 1 + wk - 1)
            for st in range(S):
                cur = np.zeros(32)
                for lane in range(w):
                    jj, kk = lane % wj, lane // wj
                    jr, kr = (jj, kk) if direction > 0 else (wj - 1 - jj, wk - 1 - kk)
                    r = st - skew * (jr + kr)
                    if not (0 <= r < nx):
                        continue
                    i = r if direction > 0 else nx - 1 - r
                    p = base + i * w + lane
                    assert tuple(ijk[p]) == (i, j0 + jj, k0 + kk)

                    def nbr_value(is_j):
                        ext = (jr == 0) if is_j else (kr == 0)
                        if not ext:
                            src_lane = lane - direction * (1 if is_j else wj)
                            return hist[st - skew][src_lane]
                        nb = nbJ if is_j else nbK
                        if nb < 0:
                            return 0.0
                        b2, w2, wj2, wk2 = tiles[nb][:4]
                        if is_j:
                            l2 = (wj2 - 1 if direction > 0 else 0) + wj2 * kk
                        else:
                            l2 = jj + wj2 * (wk2 - 1 if direction > 0 else 0)
                        v = dst[b2 + i * w2 + l2]
                        assert not np.isnan(v)
                        return v
                    vK, vJ = nbr_value(False), nbr_value(True)
                    y1 = hist[st - 1][lane] if st > 0 else 0.0
                    acc = rD[p] * src[p] if direction > 0 else src[p]
                    acc -= t[0, p] * vK
                    acc -= t[1, p] * vJ
                    acc -= t[2, p] * y1
                    cur[lane] = acc
                    assert np.isnan(dst[p])
                    dst[p] = acc
                hist[st] = cur
                hist.pop(st - skew - 1, None)
            done_tiles.add(ti)
        assert not np.isnan(dst).any()
    wA = np.empty(n)
    wA[perm] = z
    return wA


@pytest.mark.parametrize("shape", [(12, 10, 9), (5, 40, 3), (7, 6, 1), (33, 1, 1), (9, 8, 2), (4, 37, 5), (16, 17, 13)])
@pytest.mark.parametrize("sym", [True, False])
def test_pencil_schedule_reproduces_the_precondition_bit_for_bit(shape, sym, monkeypatch):
    """The tile-major layout and the pencil schedule of structured blocks (tiles in tile-wavefront order, skewed
    lanes) are a valid topological order and, with the kernel's arithmetic, reproduce DIC/DILU precondition exactly."""
    import sys as _sys
    _sys.path.insert(0, str(ROOT / "oracle"))
    import ldu_oracle as orc

    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    nx, ny, nz = shape
    s = cases.cavity_laplacian(nx, ny, nz, coeffs="random") if sym else cases.convection_diffusion(nx, ny, nz, dt_coeff=50.0)
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper)
    assert list(mesh.get_i32(21, 0)[:3]) == [nx, ny, nz]
    perm = mesh.get_i32(13, 0)
    assert np.array_equal(np.sort(perm), np.arange(s.n_cells))
    # every tile starts on an even position (16-byte aligned rows for the bulk copies) and tiles are contiguous
    tiles = mesh.get_i32(22, 0).reshape(-1, 10)
    assert (tiles[:, 0] % 2 == 0).all()
    assert np.array_equal(tiles[1:, 0], np.cumsum(nx * tiles[:-1, 1]))
    # the forward processing order handed to the wavefront kernels visits the canonical wavefronts
    fwd_pos = mesh.get_i32(20, 0)
    assert np.array_equal(perm[fwd_pos], mesh.get_i32(4, 0))
    rA = np.cos(0.37 * np.arange(s.n_cells)) + 0.1
    want = orc.precondition(orc.System(s), "DIC" if sym else "DILU", rA)
    for skew in (1, 2):
        got = _emulate_pencil_precondition(mesh, s, rA, skew)
        assert np.array_equal(got, want)


def test_no_pencil_plan_for_unstructured_addressing(monkeypatch):
    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    s = cases.random_graph(300, symmetric=True)
    mesh = capi.Mesh(s.n_cells, s.lower, s.upper)
    assert mesh.get_i32(21, 0).size == 0 and mesh.get_i32(20, 0).size == 0
    # a block with one face missing is not a block either
    lower, upper, _ = cases.block_addressing(6, 5, 4)
    mesh = capi.Mesh(120, np.delete(lower, 17), np.delete(upper, 17))
    assert mesh.get_i32(21, 0).size == 0
    # small blocks keep the wavefront-major layout unless asked, and B200LS_PENCIL=0 switches the plan off
    monkeypatch.delenv("B200LS_PENCIL_MIN_CELLS")
    m0 = capi.Mesh(120, lower, upper)
    assert m0.get_i32(21, 0).size == 0
    assert np.array_equal(m0.get_i32(13, 0), m0.get_i32(4, 0))
    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    assert capi.Mesh(120, lower, upper).get_i32(21, 0).size == 7
    monkeypatch.setenv("B200LS_PENCIL", "0")
    assert capi.Mesh(120, lower, upper).get_i32(21, 0).size == 0
