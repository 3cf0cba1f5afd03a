"""Device-resident coefficients: b200ls_matrix_set_dev takes the coefficient arrays from device
memory, b200ls_matrix_set_if_changed skips the upload (and keeps factorisation / coarse matrices) when nothing changed."""
import numpy as np
import pytest

from _util import capi, cases

pytestmark = pytest.mark.gpu


def _solve_all(capi, mat, source, symmetric):
    out = []
    for solver, kw in ((("PCG", dict(preconditioner="DIC")),) if symmetric else
                       (("PBiCGStab", dict(preconditioner="DILU")),)) + (("GAMG", dict(smoother="GaussSeidel")),):
        ctl = capi.controls(solver, tolerance=1e-10, relTol=0.0, **kw)
        psi, perf = mat.solve(ctl, source)
        out.append((psi, perf.nIterations, perf.finalResidual))
    return out


@pytest.mark.parametrize("kind", ["sym", "asym", "cyclic"])
def test_set_dev_equals_host_set(kind):
    import torch

    capi.init(0)
    if kind == "sym":
        s = cases.cavity_laplacian(14, 11, 9, coeffs="random")
    elif kind == "asym":
        s = cases.convection_diffusion(12, 10, 8)
    else:
        s = cases.add_cyclic(cases.cavity_laplacian(12, 10, 8, coeffs="random"), 0)
    lower = None if s.symmetric else s.lower_coeffs
    bou = [i.bou_coeffs for i in s.interfaces]
    inn = [i.int_coeffs for i in s.interfaces]
    mesh, mat = capi.from_system(s)
    mesh.agglomerate(s.face_weights)
    mat.set(s.diag, s.upper_coeffs, lower, bou, inn)
    want = _solve_all(capi, mat, s.source, s.symmetric)
    want_amul = mat.amul(s.source)

    # scramble the matrix, then bring the real coefficients back from device memory only
    mat.set(s.diag * 3.0, s.upper_coeffs * 0.5, None if lower is None else lower * 0.25, bou, inn)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t_d, t_u = dev(s.diag), dev(s.upper_coeffs)
    t_l = None if lower is None else dev(lower)
    t_b, t_i = [dev(b) for b in bou], [dev(b) for b in inn]
    torch.cuda.synchronize()
    mat.set_dev(t_d.data_ptr(), t_u.data_ptr(), None if t_l is None else t_l.data_ptr(),
                [t.data_ptr() for t in t_b], [t.data_ptr() for t in t_i])
    assert np.array_equal(mat.amul(s.source), want_amul)
    got = _solve_all(capi, mat, s.source, s.symmetric)
    for (p0, n0, r0), (p1, n1, r1) in zip(want, got):
        assert n0 == n1 and r0 == r1
        assert np.array_equal(p0, p1)
    mat.close()
    mesh.close()


def test_set_if_changed_skips_the_upload_and_keeps_the_hierarchy():
    capi.init(0)
    s = cases.cavity_laplacian(24, 20, 16, coeffs="random")
    mesh, mat = capi.from_system(s)
    mesh.agglomerate(s.face_weights)
    assert mat.set_if_changed(s.diag, s.upper_coeffs) is True
    ctl = capi.controls("GAMG", smoother="GaussSeidel", tolerance=1e-9, relTol=0.0)
    psi0, perf0 = mat.solve(ctl, s.source)
    assert mat.set_if_changed(s.diag.copy(), s.upper_coeffs.copy()) is False        # same values, other buffers
    psi1, perf1 = mat.solve(ctl, s.source)
    assert np.array_equal(psi0, psi1) and perf0.nIterations == perf1.nIterations
    # the coarse-level matrices were not rebuilt: fewer kernels than the first solve of these coefficients
    assert perf1.kernelLaunches < perf0.kernelLaunches
    d2 = s.diag.copy()
    d2[s.n_cells // 2] *= 1.0 + 1e-12                                               # one ulp-scale change is a change
    assert mat.set_if_changed(d2, s.upper_coeffs) is True
    assert mat.set_if_changed(d2, s.upper_coeffs, s.upper_coeffs) is True           # symmetric -> asymmetric storage
    assert mat.set_if_changed(d2, s.upper_coeffs, s.upper_coeffs) is False
    mat.set(s.diag, s.upper_coeffs)                                                  # plain set forgets the fingerprint
    assert mat.set_if_changed(s.diag, s.upper_coeffs) is True
    mat.close()
    mesh.close()
