import sys, numpy as np
sys.path.insert(0, "tests")
from _util import *
capi.init(0)
for name in FIXTURES:
    inp, ref = load_fixture(name)
    s = system_from_entries(inp)
    mesh, mat = capi.from_system(s)
    mesh.agglomerate(s.face_weights); mat.set(s.diag, s.upper_coeffs, s.lower_coeffs)
    for i, text in solve_keys(inp):
        ctl = controls_from_dict(text, recordHistory=1)
        psi, perf = mat.solve(ctl, s.source)
        r = ref[f"solve.{i}.perf"]
        h = capi.history(perf); hk=f"solve.{i}.historyResiduals"
        hd = ""
        if hk in ref:
            rh = ref[hk]; n=min(len(h),len(rh)); hd = "histRel %.1e histAbs %.1e" % (np.max(np.abs(h[:n]-rh[:n])/np.abs(rh[:n])), np.max(np.abs(h[:n]-rh[:n])))
        print(f"{name:22s} {text[:60]:60s} it {perf.nIterations:4d}/{int(r[2]):4d} init {abs(perf.initialResidual-r[0])/r[0]:.1e} finRel {abs(perf.finalResidual-r[1])/max(r[1],1e-300):.1e} psi {max_rel_diff(psi, ref[f'solve.{i}.psi']):.1e} {hd}")
