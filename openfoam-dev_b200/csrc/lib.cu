// Unity translation unit of libb200ls.so (kernels are defined once, in kernels.cuh).
#include "solver.cu"
#include "capi.cu"
