/*
 * b200ls.h -- C-ABI of libb200ls.so, the B200 (sm_100a) kernels + host logic behind
 * libB200LinearSolvers (the OpenFOAM lduMatrix solver plugin).
 *
 * Every entry point takes plain HOST pointers owned by the caller and valid only for the
 * duration of the call (unless the name ends in _dev).  Return 0 = ok, non-zero = error;
 * b200ls_last_error() gives the message, which the plugin shim turns into
 * FatalErrorInFunction (reference error convention: SURVEY.md 8(b)).
 *
 * "label" = int32_t, "scalar" = double (reference etc/bashrc:85,89 WM_LABEL_SIZE=32, WM_PRECISION_OPTION=DP).
 *
 * Reference interfaces replaced (paths relative to /root/reference/src/OpenFOAM/matrices/lduMatrix):
 *   b200ls_mesh_create     lduAddressing (lduAddressing/lduAddressing.{H,C}): lowerAddr/upperAddr +
 *                          demand-driven losort/ownerStart/losortStart (lduAddressing.C:32-170),
 *                          lduInterface::faceCells() (lduAddressing/lduInterface/lduInterface.H:82)
 *   b200ls_agglomerate     GAMGAgglomeration::New -> faceAreaPairGAMGAgglomeration ->
 *                          pairGAMGAgglomeration::agglomerate (pairGAMGAgglomerate.C:31-301) +
 *                          GAMGAgglomeration::agglomerateLduAddressing (GAMGAgglomerateLduAddressing.C:32-353)
 *   b200ls_matrix_set      lduMatrix coefficients: diag()/upper()/lower() (lduMatrix/lduMatrix.H:87,609-621) and the
 *                          interfaceBouCoeffs/interfaceIntCoeffs a solver is constructed with (lduMatrix.H:187-195)
 *   b200ls_amul/_residual/_sum_a   lduMatrix::Amul / residual / sumA (lduMatrix/lduMatrixATmul.C:34-92, 203-280, 154-200)
 *   b200ls_precondition    lduMatrix::preconditioner::precondition (lduMatrix.H:490-495):
 *                          DICPreconditioner.C:57-123, DILUPreconditioner.C:57-135
 *   b200ls_smooth          lduMatrix::smoother::smooth (lduMatrix.H:399-405):
 *                          GaussSeidelSmoother.C:66-187, DICSmoother.C:67-116, DILUSmoother.C:67-119
 *   b200ls_solve           lduMatrix::solver::solve (lduMatrix.H:250-255): PCG.C:65-193, PBiCGStab.C:68-254,
 *                          GAMGSolverSolve.C:31-145; result = solverPerformance (LduMatrix/LduMatrix/SolverPerformance.H)
 */
#ifndef B200LS_H
#define B200LS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200ls_mesh_s*   b200ls_mesh_t;
typedef struct b200ls_matrix_s* b200ls_matrix_t;

/* ---- process-level ------------------------------------------------------------------------ */

/* Select the CUDA device of this process (one process per GPU) and, when nRanks > 1, create the
 * NCCL communicator from a 128-byte ncclUniqueId obtained on rank 0 with
 * b200ls_nccl_unique_id() and broadcast by the caller (Pstream / torch.distributed / MPI).
 * Replaces UPstream::init (src/Pstream/mpi/UPstream.C:68-118) for the solver's own traffic. */
int b200ls_init(int device, const void* ncclUniqueId, int rank, int nRanks);
int b200ls_nccl_unique_id(void* out128);
void b200ls_finalize(void);
const char* b200ls_last_error(void);
/* 1 if a CUDA device is usable by this process, else 0 (no error is recorded). */
int b200ls_device_available(void);
/* number of CUDA devices visible to this process (0 without a driver): hosts map node-local ranks to devices with it */
int b200ls_device_count(void);

/* Optional host-side communicator for the once-per-mesh agglomeration of a DECOMPOSED mesh: the neighbour
 * exchange of restrictMap over each processor patch (GAMGAgglomerateLduAddressing.C:268-283) and the global sums of
 * continueAgglomerating (GAMGAgglomeration.C:211-229).  exchange(): send[i] (sizes[i] labels) goes to rank nbr[i],
 * recv[i] receives as many from it.  Without it (default) the library uses NCCL on small staging buffers.
 * Pass NULL callbacks to unset. */
typedef void (*b200ls_exchange_fn)(int32_t nIfaces, const int32_t* nbr, const int32_t* sizes,
                                   const int32_t* const* send, int32_t* const* recv);
typedef int64_t (*b200ls_sum_fn)(int64_t localValue);
int b200ls_set_host_comm(int32_t rank, int32_t nRanks, b200ls_exchange_fn exchange, b200ls_sum_fn sum);

/* ---- mesh (cached per lduAddressing by the caller) ---------------------------------------------- */

/* Host analysis only (no device work): losort/ownerStart/losortStart, forward/backward wavefronts,
 * the row permutation (wavefront-major, or tile-major on structured blocks) and the native row layout.  Interfaces: one entry per COUPLED patch, in patch order.
 * ifaceNeighbRank[i] >= 0: processor patch to that rank (processorLduInterface::neighbProcNo()).
 * ifaceNeighbRank[i] <  0: one half of a cyclic pair on this rank (lduAddressing/lduInterface/cyclicLduInterface.H):
 *   the value is B200LS_CYCLIC(p) = -(1 + p), p = cyclicLduInterface::nbrPatchIndex() counted among the coupled
 *   patches passed here; face e of patch i is coupled to face e of patch p, and the half with the lower index is the
 *   owner.  Scalar (rank-0, untransformed) coupling only. */
#define B200LS_CYCLIC(partnerPatch) (-(1 + (partnerPatch)))
b200ls_mesh_t b200ls_mesh_create(int32_t nCells, int32_t nFaces,
                                 const int32_t* lower, const int32_t* upper,
                                 int32_t nInterfaces, const int32_t* ifaceSizes,
                                 const int32_t* const* ifaceFaceCells,
                                 const int32_t* ifaceNeighbRank);
void b200ls_mesh_free(b200ls_mesh_t mesh);

enum b200ls_i32_which {
    B200LS_LOSORT = 0,              /* lduAddressing::losortAddr()                       */
    B200LS_OWNER_START = 1,         /* lduAddressing::ownerStartAddr()                   */
    B200LS_LOSORT_START = 2,        /* lduAddressing::losortStartAddr()                  */
    B200LS_FWD_LEVEL_OFFSETS = 3,   /* canonical forward wavefront offsets [nLevels+1]   */
    B200LS_FWD_LEVEL_ROWS = 4,      /* cells of each forward wavefront, ascending        */
    B200LS_BWD_LEVEL_OFFSETS = 5,
    B200LS_BWD_LEVEL_ROWS = 6,
    B200LS_RESTRICT_ADDRESSING = 7, /* GAMGAgglomeration::restrictAddressing(level)      */
    B200LS_FACE_RESTRICT_ADDRESSING = 8, /* ::faceRestrictAddressing(level)              */
    B200LS_FACE_FLIP_MAP = 9,       /* ::faceFlipMap(level) widened to int32 0/1         */
    B200LS_LOWER_ADDR = 10,         /* meshLevel(level).lduAddr().lowerAddr()            */
    B200LS_UPPER_ADDR = 11,         /* meshLevel(level).lduAddr().upperAddr()            */
    B200LS_LEVEL_SIZES = 12,        /* {nCells, nFaces} of meshLevel(level)              */
    /* native device layout, exposed for the host-logic tests: rows live at "positions" */
    B200LS_PERM = 13,               /* position -> cell (forward-wavefront-major, or tile-major: see below) */
    B200LS_LPTR = 14,               /* CSR of the neighbour-side (lower) triangle, by position */
    B200LS_LCOL = 15,               /*   column = position of the coupled row            */
    B200LS_LFACE = 16,              /*   face of each entry                              */
    B200LS_UPTR = 17,               /* CSR of the owner-side (upper) triangle            */
    B200LS_UCOL = 18,
    B200LS_UFACE = 19,
    /* structured blocks (an nx*ny*nz hex block numbered i-fastest) are laid out tile-major for the pencil sweeps
     * (DESIGN.md 3.2); all three are empty when the level keeps the wavefront-major layout */
    B200LS_FWD_POS = 20,            /* forward processing order (wavefront by wavefront) -> position     */
    B200LS_PENCIL_DIMS = 21,        /* {nx, ny, nz, WJ, WK, nJ, nK}                                      */
    B200LS_PENCIL_TILES = 22,       /* per tile (memory order, J fastest): {base, w, wj, wk, j0, k0,
                                       tile index of (J-1,K), (J,K-1), (J+1,K), (J,K+1) or -1}          */
    B200LS_PENCIL_ORDER = 23        /* launch order of the forward sweeps (tile wavefronts)              */
};
/* level 0 = the finest mesh; level k>0 = k-th coarse mesh (= reference meshLevel(k)).
 * RESTRICT_ADDRESSING / FACE_RESTRICT_ADDRESSING / FACE_FLIP_MAP at `level` map level -> level+1,
 * as in the reference.  The pointer stays valid until the mesh is freed or re-agglomerated. */
int b200ls_mesh_get_i32(b200ls_mesh_t mesh, int which, int level, const int32_t** data, int64_t* n);
int b200ls_mesh_n_levels(b200ls_mesh_t mesh);

enum b200ls_iface_i32_which {
    B200LS_IFACE_FACE_CELLS = 0,               /* lduInterface::faceCells() of coupled patch `iface` at `level`   */
    B200LS_IFACE_FACE_RESTRICT_ADDRESSING = 1  /* GAMGInterface::faceRestrictAddressing(): level -> level+1       */
};
int b200ls_mesh_get_iface_i32(b200ls_mesh_t mesh, int which, int level, int iface, const int32_t** data,
                              int64_t* n);   /* number of mesh levels incl. the finest */

/* Pair agglomeration with the given finest-level face weights (faceAreaPair passes
 * mag(cmptMultiply(Sf/sqrt(magSf), (1 1.01 1.02)))).  minCellsPerProcessor: reference default 10
 * (GAMGAgglomeration.C:250-257); mergeLevels: default 1 (pairGAMGAgglomeration.C:46);
 * forwardStart: initial value of the reference's process-global pairGAMGAgglomeration::forward_
 * (true in a fresh process).  Returns the number of coarse levels created, <0 on error. */
int b200ls_agglomerate(b200ls_mesh_t mesh, const double* faceWeights,
                       int32_t minCellsPerProcessor, int32_t mergeLevels, int32_t forwardStart);

/* Same, with the cell maps of every level supplied by the caller: restrictAddr[k] maps the cells of level k to
 * the nCoarseCells[k] cells of level k+1 (GAMGAgglomeration::restrictAddressing(k) / nCells(k)).  The plugin uses
 * this with the reference's own cached GAMGAgglomeration MeshObject (GAMGAgglomeration.C:349-400). */
int b200ls_agglomerate_from_maps(b200ls_mesh_t mesh, int32_t nCoarseLevels, const int32_t* const* restrictAddr,
                                 const int32_t* nCoarseCells);

/* ---- matrix --------------------------------------------------------------------------------------------- */

b200ls_matrix_t b200ls_matrix_create(b200ls_mesh_t mesh);
void b200ls_matrix_free(b200ls_matrix_t m);
/* Upload coefficients (H2D) and re-lay them out in the native format.  lower == NULL => symmetric.
 * ifaceBouCoeffs/ifaceIntCoeffs: one array per coupled patch (may be NULL when nInterfaces == 0). */
int b200ls_matrix_set(b200ls_matrix_t m, const double* diag, const double* upper, const double* lower,
                      const double* const* ifaceBouCoeffs, const double* const* ifaceIntCoeffs);

/* The same with DEVICE pointers (diag/upper/lower and every bou[i]/inn[i] live on this GPU, reference cell / face
 * order; the pointer arrays themselves are host memory): the coefficients of a GPU-side assembly never visit the host.
 * Replaces the host copies of fvMatrix::solveSegregated's arguments (fvMatrixSolve.C:150-190). */
int b200ls_matrix_set_dev(b200ls_matrix_t m, const double* diag, const double* upper, const double* lower,
                          const double* const* ifaceBouCoeffs, const double* const* ifaceIntCoeffs);
/* b200ls_matrix_set that first compares a 64-bit fingerprint of all arrays with the one of the coefficients the
 * matrix already holds: when equal nothing is copied and the factorisation / coarse-level matrices stay valid
 * (*changed = 0).  PISO solves the same pressure matrix once per corrector with a new source only
 * (pEqn.H of icoFoam: the matrix is rebuilt from the same rAU). */
int b200ls_matrix_set_if_changed(b200ls_matrix_t m, const double* diag, const double* upper, const double* lower,
                                 const double* const* ifaceBouCoeffs, const double* const* ifaceIntCoeffs,
                                 int32_t* changed);

int b200ls_amul(b200ls_matrix_t m, const double* psi, double* Apsi);
int b200ls_residual(b200ls_matrix_t m, const double* psi, const double* source, double* rA);
int b200ls_sum_a(b200ls_matrix_t m, double* sumA);

enum b200ls_solver {
    B200LS_PCG = 0, B200LS_PBICGSTAB = 1, B200LS_GAMG = 2, B200LS_SMOOTH_SOLVER = 3,
    B200LS_DIAGONAL_SOLVER = 4      /* solvers/diagonalSolver/diagonalSolver.C:62-79: psi = source/diag, 0 iterations */
};
enum b200ls_precond {            /* preconditioner (Krylov) or smoother (GAMG / smoothSolver) */
    B200LS_NONE = 0, B200LS_DIAGONAL = 1, B200LS_DIC = 2, B200LS_DILU = 3, B200LS_GAUSS_SEIDEL = 4,
    B200LS_SYM_GAUSS_SEIDEL = 5,    /* smoothers/symGaussSeidel/symGaussSeidelSmoother.C:66-217                  */
    B200LS_DIC_GAUSS_SEIDEL = 6,    /* smoothers/DICGaussSeidel/DICGaussSeidelSmoother.C:79-89                   */
    B200LS_DILU_GAUSS_SEIDEL = 7,   /* smoothers/DILUGaussSeidel/DILUGaussSeidelSmoother.C                       */
    B200LS_GAMG_PRECOND = 8         /* preconditioners/GAMGPreconditioner/GAMGPreconditioner.C:81-148            */
};

int b200ls_precondition(b200ls_matrix_t m, int precond, const double* rA, double* wA);
/* DIC/DILU reciprocal diagonal (DICPreconditioner::calcReciprocalD) in cell order. */
int b200ls_reciprocal_d(b200ls_matrix_t m, int precond, double* rD);
int b200ls_smooth(b200ls_matrix_t m, int smoother, double* psi, const double* source, int32_t nSweeps);

typedef struct b200ls_controls {
    int32_t solver;                 /* enum b200ls_solver                                     */
    int32_t precond;                /* enum b200ls_precond: preconditioner or smoother       */
    double  tolerance;              /* lduMatrixSolver.C:158-164 defaults 1e-6 / 0 / 1000 / 0 */
    double  relTol;
    int32_t maxIter;
    int32_t minIter;
    /* GAMG (GAMGSolver.C:70-80 defaults in brackets) */
    int32_t nPreSweeps;             /* [0] */
    int32_t preSweepsLevelMultiplier;  /* [1] */
    int32_t maxPreSweeps;           /* [4] */
    int32_t nPostSweeps;            /* [2] */
    int32_t postSweepsLevelMultiplier; /* [1] */
    int32_t maxPostSweeps;          /* [4] */
    int32_t nFinestSweeps;          /* [2] */
    int32_t scaleCorrection;        /* [-1 = matrix.symmetric()] 0/1 */
    int32_t nSweeps;                /* smoothSolver [1] */
    int32_t recordHistory;          /* store the residual after every iteration in perf.history */
    /* preconditioner GAMG (sub-dictionary `preconditioner { preconditioner GAMG; smoother ...; nVcycles 2; }`):
     * the V-cycle controls above apply to it, plus: */
    int32_t precSmoother;           /* enum b200ls_precond: smoother of the preconditioning V-cycles [GaussSeidel] */
    int32_t nVcycles;               /* [2] */
    double  precTolerance;          /* tolerance / relTol of the sub-dictionary: inherited by its coarsest solver */
    double  precRelTol;
} b200ls_controls;

void b200ls_controls_default(b200ls_controls* c);

#define B200LS_MAX_HISTORY 4096

typedef struct b200ls_perf {        /* SolverPerformance<scalar> + timing */
    double  initialResidual;
    double  finalResidual;
    int32_t nIterations;
    int32_t converged;
    int32_t singular;
    int32_t nHistory;
    double  normFactor;
    double  solveMs;                /* device time of the solve loop (CUDA events), no H2D/D2H */
    double  setupMs;                /* per-solve setup on device: factorisation, coarse matrices */
    double  h2dMs;                  /* psi/source upload + psi download                       */
    int64_t kernelLaunches;         /* kernels launched by this call                          */
    double  history[B200LS_MAX_HISTORY];
} b200ls_perf;

/* psi: in = initial guess, out = solution.  Equivalent of
 * lduMatrix::solver::New(name, matrix, bouCoeffs, intCoeffs, interfaces, dict)->solve(psi, source). */
int b200ls_solve(b200ls_matrix_t m, const b200ls_controls* c, double* psi, const double* source,
                 b200ls_perf* perf);

/* Same with psi/source already resident in device memory (cell order, device pointers). */
int b200ls_solve_dev(b200ls_matrix_t m, const b200ls_controls* c, double* psi_dev, const double* source_dev,
                     b200ls_perf* perf);

/* ---- instrumentation -------------------------------------------------------------------------------- */

/* Time `reps` launches of one kernel class on the matrix with CUDA events on the launching stream.
 * which: 0 = Amul, 1 = DIC/DILU precondition, 2 = GaussSeidel sweep, 3 = PCG vector ops of one iteration.
 * Returns average milliseconds per launch in *ms. */
int b200ls_time_kernel(b200ls_matrix_t m, int which, int reps, double* ms);

#ifdef __cplusplus
}
#endif

#endif /* B200LS_H */
