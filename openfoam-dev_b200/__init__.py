"""libB200LinearSolvers: B200-native lduMatrix solver stack (PCG / PBiCGStab / GAMG).

The product is the C-ABI library built from csrc/ (include/b200ls.h) and the OpenFOAM
plugin shim in plugin/.  The Python modules here are test/bench glue only:
  ldu_io  - B2LS container IO shared with the oracle harness
  cases   - synthetic LDU systems of BASELINE.json's configurations
  capi    - ctypes binding of include/b200ls.h
  build   - nvcc/g++ build recipes
"""
