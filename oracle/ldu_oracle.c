/*
 * oracle/ldu_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (openfoam-dev_b200/libb200ls.so); used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * A plain-C, single-threaded restatement of the reference's scalar lduMatrix solver stack in the reference's own
 * face-loop (scatter) formulation and sequential summation order.  Paths below are relative to
 * /root/reference/src/OpenFOAM/matrices/lduMatrix.  Compile with -ffp-contract=off (no FMA, like the reference's
 * x86-64 build).  PARITY PINNED: tests/test_oracle.py checks every function here against tests/golden/*.b2ls,
 * which hold outputs of the unmodified reference run by oracle/ref_harness.C.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* A coupled patch whose neighbour cells live in the same address space: one half of a cyclic pair
 * (lduAddressing/lduInterface/cyclicLduInterface.H; face e is coupled to face e of the neighbour patch, so
 * nbrCells = faceCells of that patch).  Processor patches are emulated on top of this file by tests/_emulated_ranks.py. */
typedef struct {
    int32_t n;
    const int32_t *faceCells, *nbrCells;
    const double *bou, *inn;       /* interfaceBouCoeffs, interfaceIntCoeffs */
} iface_t;

typedef struct {
    int32_t nCells, nFaces;
    const int32_t *l, *u;          /* lowerAddr, upperAddr */
    const double *diag, *upper, *lower; /* lower == upper when symmetric */
    int32_t nIfaces;
    const iface_t* ifaces;
} ldu_t;

typedef struct {
    double initialResidual, finalResidual, normFactor;
    int32_t nIterations, converged, singular, nHistory;
    double history[4096];
} perf_t;

typedef struct {
    double tolerance, relTol;
    int32_t maxIter, minIter;
    int32_t precond;               /* 0 none, 1 diagonal, 2 DIC, 3 DILU, 4 GaussSeidel */
    int32_t nSweeps;
    int32_t nPreSweeps, preSweepsLevelMultiplier, maxPreSweeps;
    int32_t nPostSweeps, postSweepsLevelMultiplier, maxPostSweeps, nFinestSweeps, scaleCorrection;
    int32_t precSmoother, nVcycles;   /* preconditioner GAMG (kind 8): smoother and cycles of the V-cycles */
    double precTolerance, precRelTol; /* ... and the tolerances its coarsest-level solver inherits */
    void* hierarchy;               /* hierarchy_t* for preconditioner GAMG */
} ctl_t;

static const double VSMALL = 2.2250738585072014e-308; /* vSmall, primitives/Scalar/doubleScalar/doubleScalar.H:57 */
static const double SMALL_ = 1e-20;                    /* solverPerformance::small_, SolverPerformance.H:288 */

/* ---- lduMatrix::Amul / residual / sumA  (lduMatrix/lduMatrixATmul.C:34-92, 203-280, 154-200) ---- */

/* lduMatrix::updateMatrixInterfaces (lduMatrix/lduMatrixUpdateMatrixInterfaces.C:94-266), patches in index order, with
 * the cyclic update  result[faceCells[e]] -= coeffs[e]*psi[nbrFaceCells[e]]  (finiteVolume cyclicFvPatchField.C:150-174,
 * solvers/GAMG/interfaceFields/cyclicGAMGInterfaceField/cyclicGAMGInterfaceField.C:124-146).  negate: the caller passes
 * -interfaceBouCoeffs (residual, lduMatrixATmul.C:236-244; GaussSeidelSmoother.C:95-107). */
static void update_interfaces(const ldu_t* A, int negate, const double* psi, double* result) {
    for (int i = 0; i < A->nIfaces; i++) {
        const iface_t* I = &A->ifaces[i];
        for (int e = 0; e < I->n; e++) {
            const double c = negate ? -I->bou[e] : I->bou[e];
            result[I->faceCells[e]] -= c * psi[I->nbrCells[e]];
        }
    }
}

void oracle_amul(const ldu_t* A, const double* psi, double* Apsi) {
    for (int c = 0; c < A->nCells; c++) Apsi[c] = A->diag[c] * psi[c];
    for (int f = 0; f < A->nFaces; f++) {
        Apsi[A->u[f]] += A->lower[f] * psi[A->l[f]];
        Apsi[A->l[f]] += A->upper[f] * psi[A->u[f]];
    }
    update_interfaces(A, 0, psi, Apsi);
}

void oracle_residual(const ldu_t* A, const double* psi, const double* source, double* rA) {
    for (int c = 0; c < A->nCells; c++) rA[c] = source[c] - A->diag[c] * psi[c];
    for (int f = 0; f < A->nFaces; f++) {
        rA[A->u[f]] -= A->lower[f] * psi[A->l[f]];
        rA[A->l[f]] -= A->upper[f] * psi[A->u[f]];
    }
    update_interfaces(A, 1, psi, rA);
}

void oracle_sum_a(const ldu_t* A, double* sumA) {
    for (int c = 0; c < A->nCells; c++) sumA[c] = A->diag[c];
    for (int f = 0; f < A->nFaces; f++) {
        sumA[A->u[f]] += A->lower[f];
        sumA[A->l[f]] += A->upper[f];
    }
    for (int i = 0; i < A->nIfaces; i++)                   /* lduMatrixATmul.C:185-199 */
        for (int e = 0; e < A->ifaces[i].n; e++) sumA[A->ifaces[i].faceCells[e]] -= A->ifaces[i].bou[e];
}

/* ---- lduAddressing::calcLosort (lduAddressing/lduAddressing.C:32-90) ---- */

static int32_t* make_losort(const ldu_t* A) {
    int32_t* start = calloc((size_t)A->nCells + 1, sizeof(int32_t));
    int32_t* lst = malloc(sizeof(int32_t) * (size_t)(A->nFaces > 0 ? A->nFaces : 1));
    for (int f = 0; f < A->nFaces; f++) start[A->u[f] + 1]++;
    for (int c = 0; c < A->nCells; c++) start[c + 1] += start[c];
    for (int f = 0; f < A->nFaces; f++) lst[start[A->u[f]]++] = f;
    free(start);
    return lst;
}

void oracle_losort(const ldu_t* A, int32_t* out) {
    int32_t* l = make_losort(A);
    memcpy(out, l, sizeof(int32_t) * (size_t)A->nFaces);
    free(l);
}

/* ---- DIC / DILU  (preconditioners/DICPreconditioner/DICPreconditioner.C:57-123,
 *                   preconditioners/DILUPreconditioner/DILUPreconditioner.C:57-135) ---- */

void oracle_calc_reciprocal_d(const ldu_t* A, double* rD) {
    for (int c = 0; c < A->nCells; c++) rD[c] = A->diag[c];
    for (int f = 0; f < A->nFaces; f++) rD[A->u[f]] -= A->upper[f] * A->lower[f] / rD[A->l[f]];
    for (int c = 0; c < A->nCells; c++) rD[c] = 1.0 / rD[c];
}

/* DIC walks the faces in face order, DILU in losort order: pass losort = NULL for DIC */
static void precondition_sweeps(const ldu_t* A, const double* rD, const int32_t* losort, double* wA) {
    for (int f = 0; f < A->nFaces; f++) {
        const int s = losort ? losort[f] : f;
        wA[A->u[s]] -= rD[A->u[s]] * A->lower[s] * wA[A->l[s]];
    }
    for (int f = A->nFaces - 1; f >= 0; f--) wA[A->l[f]] -= rD[A->l[f]] * A->upper[f] * wA[A->u[f]];
}

void oracle_precondition(const ldu_t* A, int kind, const double* rD, const double* rA, double* wA) {
    if (kind == 0) {                                       /* noPreconditioner */
        for (int c = 0; c < A->nCells; c++) wA[c] = rA[c];
        return;
    }
    for (int c = 0; c < A->nCells; c++) wA[c] = rD[c] * rA[c];
    if (kind == 1) return;                                 /* diagonalPreconditioner: rD = 1/diag */
    int32_t* losort = kind == 3 ? make_losort(A) : NULL;
    precondition_sweeps(A, rD, losort, wA);
    free(losort);
}

static void make_rD(const ldu_t* A, int kind, double* rD) {
    if (kind == 1) {
        for (int c = 0; c < A->nCells; c++) rD[c] = 1.0 / A->diag[c];
    } else if (kind == 2 || kind == 3) {
        oracle_calc_reciprocal_d(A, rD);
    }
}

/* ---- smoothers (smoothers/GaussSeidel/GaussSeidelSmoother.C:66-187, smoothers/DIC/DICSmoother.C:67-116,
 *                 smoothers/DILU/DILUSmoother.C:67-119) ---- */

static int32_t* make_owner_start(const ldu_t* A) {
    int32_t* os = calloc((size_t)A->nCells + 1, sizeof(int32_t));
    for (int f = 0; f < A->nFaces; f++) os[A->l[f] + 1]++;
    for (int c = 0; c < A->nCells; c++) os[c + 1] += os[c];
    return os;
}

void oracle_smooth(const ldu_t* A, int kind, double* psi, const double* source, int nSweeps) {
    const int n = A->nCells;
    if (kind == 6 || kind == 7) {                          /* DICGaussSeidel / DILUGaussSeidel: DICGaussSeidelSmoother.C:79-89 */
        oracle_smooth(A, kind == 6 ? 2 : 3, psi, source, nSweeps);
        oracle_smooth(A, 4, psi, source, nSweeps);
        return;
    }
    if (kind == 5) {                                       /* symGaussSeidelSmoother.C:125-206 */
        int32_t* os = make_owner_start(A);
        double* bPrime = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        for (int sweep = 0; sweep < nSweeps; sweep++) {
            memcpy(bPrime, source, sizeof(double) * (size_t)n);
            update_interfaces(A, 1, psi, bPrime);
            for (int c = 0; c < n; c++) {
                double psii = bPrime[c];
                for (int f = os[c]; f < os[c + 1]; f++) psii -= A->upper[f] * psi[A->u[f]];
                psii /= A->diag[c];
                for (int f = os[c]; f < os[c + 1]; f++) bPrime[A->u[f]] -= A->lower[f] * psii;
                psi[c] = psii;
            }
            for (int c = n - 1; c >= 0; c--) {
                double psii = bPrime[c];
                for (int f = os[c]; f < os[c + 1]; f++) psii -= A->upper[f] * psi[A->u[f]];
                psii /= A->diag[c];
                for (int f = os[c]; f < os[c + 1]; f++) bPrime[A->u[f]] -= A->lower[f] * psii;
                psi[c] = psii;
            }
        }
        free(bPrime);
        free(os);
        return;
    }
    if (kind == 4) {
        int32_t* os = make_owner_start(A);
        double* bPrime = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        for (int sweep = 0; sweep < nSweeps; sweep++) {
            memcpy(bPrime, source, sizeof(double) * (size_t)n);
            update_interfaces(A, 1, psi, bPrime);
            for (int c = 0; c < n; c++) {
                double psii = bPrime[c];
                for (int f = os[c]; f < os[c + 1]; f++) psii -= A->upper[f] * psi[A->u[f]];
                psii /= A->diag[c];
                for (int f = os[c]; f < os[c + 1]; f++) bPrime[A->u[f]] -= A->lower[f] * psii;
                psi[c] = psii;
            }
        }
        free(bPrime);
        free(os);
        return;
    }
    /* DIC / DILU smoother: rA = residual; rA *= rD; forward/backward in plain face order; psi += rA */
    double* rD = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double* rA = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    oracle_calc_reciprocal_d(A, rD);
    for (int sweep = 0; sweep < nSweeps; sweep++) {
        oracle_residual(A, psi, source, rA);
        for (int c = 0; c < n; c++) rA[c] *= rD[c];
        precondition_sweeps(A, rD, NULL, rA);
        for (int c = 0; c < n; c++) psi[c] += rA[c];
    }
    free(rD);
    free(rA);
}

/* ---- reductions (fields/Field/FieldReductionFunctions.C:190-292): sequential left-to-right ---- */

static double sum_prod(const double* a, const double* b, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
static double sum_mag(const double* a, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) s += fabs(a[i]);
    return s;
}

/* ---- lduMatrix::solver::normFactor (lduMatrix/lduMatrixSolver.C:174-197) ---- */

double oracle_norm_factor(const ldu_t* A, const double* psi, const double* source, const double* Apsi, double* tmp) {
    const int n = A->nCells;
    oracle_sum_a(A, tmp);
    double s = 0;
    for (int i = 0; i < n; i++) s += psi[i];
    const double avg = s / n;                              /* gAverage */
    for (int i = 0; i < n; i++) tmp[i] *= avg;
    double nf = 0;
    for (int i = 0; i < n; i++) nf += fabs(Apsi[i] - tmp[i]) + fabs(source[i] - tmp[i]);
    return nf + SMALL_;
}

/* SolverPerformance::checkConvergence (LduMatrix/LduMatrix/SolverPerformance.C:60-92): the converged flag a solve
 * reports is the result of the LAST CALL the loop conditions actually made -- when maxIter is reached the `&&`
 * short-circuits the call, so the flag keeps the value of the previous iteration's check. */
static int check(perf_t* p, double tol, double relTol) {
    p->converged = p->finalResidual < tol || (relTol > SMALL_ && p->finalResidual < relTol * p->initialResidual);
    return p->converged;
}

static void record(perf_t* p) {
    if (p->nHistory < 4096) p->history[p->nHistory++] = p->finalResidual;
}

static void gamg_precondition(const ctl_t* c, const double* rA, double* wA);

static void apply_precond(const ldu_t* A, const ctl_t* c, const double* rD, const double* rA, double* wA) {
    if (c->precond == 8) gamg_precondition(c, rA, wA);
    else oracle_precondition(A, c->precond, rD, rA, wA);
}

/* ---- PCG (solvers/PCG/PCG.C:65-193) ---- */

void oracle_pcg(const ldu_t* A, const ctl_t* c, double* psi, const double* source, perf_t* perf) {
    const int n = A->nCells;
    double* pA = malloc(sizeof(double) * (size_t)n * 4);
    double *wA = pA + n, *rA = wA + n, *rD = rA + n;
    memset(perf, 0, sizeof(*perf));
    double wArA = 1e20, wArAold;                           /* solverPerf.great_ */
    oracle_amul(A, psi, wA);
    for (int i = 0; i < n; i++) rA[i] = source[i] - wA[i];
    const double nf = oracle_norm_factor(A, psi, source, wA, pA);
    perf->normFactor = nf;
    perf->initialResidual = sum_mag(rA, n) / nf;
    perf->finalResidual = perf->initialResidual;
    if (c->minIter > 0 || !check(perf, c->tolerance, c->relTol)) {
        make_rD(A, c->precond, rD);
        do {
            wArAold = wArA;
            apply_precond(A, c, rD, rA, wA);
            wArA = sum_prod(wA, rA, n);
            if (perf->nIterations == 0) {
                for (int i = 0; i < n; i++) pA[i] = wA[i];
            } else {
                const double beta = wArA / wArAold;
                for (int i = 0; i < n; i++) pA[i] = wA[i] + beta * pA[i];
            }
            oracle_amul(A, pA, wA);
            const double wApA = sum_prod(wA, pA, n);
            if (fabs(wApA) / nf < VSMALL) {
                perf->singular = 1;
                break;
            }
            const double alpha = wArA / wApA;
            for (int i = 0; i < n; i++) {
                psi[i] += alpha * pA[i];
                rA[i] -= alpha * wA[i];
            }
            perf->finalResidual = sum_mag(rA, n) / nf;
            record(perf);
        } while ((++perf->nIterations < c->maxIter && !check(perf, c->tolerance, c->relTol)) ||
                 perf->nIterations < c->minIter);
    }
    free(pA);
}

/* ---- PBiCGStab (solvers/PBiCGStab/PBiCGStab.C:68-254) ---- */

void oracle_pbicgstab(const ldu_t* A, const ctl_t* c, double* psi, const double* source, perf_t* perf) {
    const int n = A->nCells;
    double* pA = malloc(sizeof(double) * (size_t)n * 10);
    double *yA = pA + n, *rA = yA + n, *AyA = rA + n, *sA = AyA + n, *zA = sA + n, *tA = zA + n, *rA0 = tA + n,
           *rD = rA0 + n;
    memset(perf, 0, sizeof(*perf));
    oracle_amul(A, psi, yA);
    for (int i = 0; i < n; i++) rA[i] = source[i] - yA[i];
    const double nf = oracle_norm_factor(A, psi, source, yA, pA);
    perf->normFactor = nf;
    perf->initialResidual = sum_mag(rA, n) / nf;
    perf->finalResidual = perf->initialResidual;
    if (c->minIter > 0 || !check(perf, c->tolerance, c->relTol)) {
        memcpy(rA0, rA, sizeof(double) * (size_t)n);
        double rA0rA = 0, alpha = 0, omega = 0;
        make_rD(A, c->precond, rD);
        do {
            const double rA0rAold = rA0rA;
            rA0rA = sum_prod(rA0, rA, n);
            if (fabs(rA0rA) < VSMALL) {
                perf->singular = 1;
                break;
            }
            if (perf->nIterations == 0) {
                for (int i = 0; i < n; i++) pA[i] = rA[i];
            } else {
                if (fabs(omega) < VSMALL) {
                    perf->singular = 1;
                    break;
                }
                const double beta = (rA0rA / rA0rAold) * (alpha / omega);
                for (int i = 0; i < n; i++) pA[i] = rA[i] + beta * (pA[i] - omega * AyA[i]);
            }
            apply_precond(A, c, rD, pA, yA);
            oracle_amul(A, yA, AyA);
            const double rA0AyA = sum_prod(rA0, AyA, n);
            alpha = rA0rA / rA0AyA;
            for (int i = 0; i < n; i++) sA[i] = rA[i] - alpha * AyA[i];
            perf->finalResidual = sum_mag(sA, n) / nf;
            if (++perf->nIterations >= c->minIter && check(perf, c->tolerance, c->relTol)) {
                for (int i = 0; i < n; i++) psi[i] += alpha * yA[i];
                record(perf);
                perf->converged = 1;
                free(pA);
                return;
            }
            apply_precond(A, c, rD, sA, zA);
            oracle_amul(A, zA, tA);
            double tAtA = 0;
            for (int i = 0; i < n; i++) tAtA += tA[i] * tA[i];
            omega = sum_prod(tA, sA, n) / tAtA;
            for (int i = 0; i < n; i++) {
                psi[i] += alpha * yA[i] + omega * zA[i];
                rA[i] = sA[i] - omega * tA[i];
            }
            perf->finalResidual = sum_mag(rA, n) / nf;
            record(perf);
        } while ((perf->nIterations < c->maxIter && !check(perf, c->tolerance, c->relTol)) ||
                 perf->nIterations < c->minIter);
    }
    free(pA);
}

/* ---- smoothSolver (solvers/smoothSolver/smoothSolver.C:77-193), nSweeps >= 0 ---- */

void oracle_smooth_solver(const ldu_t* A, const ctl_t* c, double* psi, const double* source, perf_t* perf) {
    const int n = A->nCells;
    double* Apsi = malloc(sizeof(double) * (size_t)n * 2);
    double* tmp = Apsi + n;
    memset(perf, 0, sizeof(*perf));
    oracle_amul(A, psi, Apsi);
    const double nf = oracle_norm_factor(A, psi, source, Apsi, tmp);
    perf->normFactor = nf;
    double s = 0;
    for (int i = 0; i < n; i++) s += fabs(source[i] - Apsi[i]);
    perf->initialResidual = s / nf;
    perf->finalResidual = perf->initialResidual;
    if (c->minIter > 0 || !check(perf, c->tolerance, c->relTol)) {
        do {
            oracle_smooth(A, c->precond, psi, source, c->nSweeps);
            oracle_residual(A, psi, source, tmp);
            perf->finalResidual = sum_mag(tmp, n) / nf;
            record(perf);
        } while (((perf->nIterations += c->nSweeps) < c->maxIter && !check(perf, c->tolerance, c->relTol)) ||
                 perf->nIterations < c->minIter);
    }
    free(Apsi);
}

/* ======================================================================================================
 * GAMG: pair agglomeration (solvers/GAMG/GAMGAgglomerations/pairGAMGAgglomeration/pairGAMGAgglomerate.C:31-301),
 * coarse addressing (GAMGAgglomeration/GAMGAgglomerateLduAddressing.C:32-353), coarse matrices
 * (solvers/GAMG/GAMGSolverAgglomerateMatrix.C:33-193), V-cycle (GAMGSolverSolve.C:31-552), scale
 * (GAMGSolverScale.C:31-76)
 * ====================================================================================================== */

typedef struct {
    int32_t nCells, nFaces;
    int32_t *l, *u;
    double *diag, *upper, *lower;  /* lower aliases upper when symmetric */
    int32_t* restrictAddr;         /* to the next level (NULL on the coarsest) */
    int32_t* faceRestrictAddr;
    uint8_t* faceFlip;
    double *corr, *src;
    int32_t nIfaces;
    struct lev_iface* ifs;         /* cyclic patches of this level */
    iface_t* views;                /* the same as ldu_t views */
} level_t;

typedef struct lev_iface {
    int32_t n, nbrPatch;
    int32_t *faceCells, *nbrCells;
    int32_t* faceRestrict;         /* patch face -> coarse patch face (NULL on the coarsest level) */
    double *bou, *inn;
} lev_iface_t;

typedef struct {
    int nLevels;                   /* incl. the finest */
    level_t lev[52];
    int symmetric;
} hierarchy_t;

static int32_t* pair_agglomerate(int32_t* nCoarseOut, int n, int nF, const int32_t* l, const int32_t* u,
                                 const double* w, int* forward) {
    /* cellFaces: neighbour-side faces first (ascending), then owner-side faces (ascending) */
    int32_t* offs = calloc((size_t)n + 1, sizeof(int32_t));
    int32_t* cnt = calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    int32_t* cf = malloc(sizeof(int32_t) * (size_t)(2 * nF > 0 ? 2 * nF : 1));
    for (int f = 0; f < nF; f++) { offs[u[f] + 1]++; offs[l[f] + 1]++; }
    for (int c = 0; c < n; c++) offs[c + 1] += offs[c];
    for (int f = 0; f < nF; f++) cf[offs[u[f]] + cnt[u[f]]++] = f;
    for (int f = 0; f < nF; f++) cf[offs[l[f]] + cnt[l[f]]++] = f;
    int32_t* map = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int c = 0; c < n; c++) map[c] = -1;
    const double great = 1.0 / 2.220446049250313e-16;
    int nCoarse = 0;
    for (int ci = 0; ci < n; ci++) {
        const int c = *forward ? ci : n - ci - 1;
        if (map[c] >= 0) continue;
        int match = -1;
        double best = -great;
        for (int k = offs[c]; k < offs[c + 1]; k++) {
            const int f = cf[k];
            if (map[u[f]] < 0 && map[l[f]] < 0 && w[f] > best) { match = f; best = w[f]; }
        }
        if (match >= 0) {
            map[u[match]] = nCoarse;
            map[l[match]] = nCoarse;
            nCoarse++;
        } else {
            int join = -1;
            best = -great;
            for (int k = offs[c]; k < offs[c + 1]; k++) {
                const int f = cf[k];
                if (w[f] > best) { join = f; best = w[f]; }
            }
            if (join >= 0) map[c] = map[u[join]] > map[l[join]] ? map[u[join]] : map[l[join]];
        }
    }
    for (int ci = 0; ci < n; ci++) {
        const int c = *forward ? ci : n - ci - 1;
        if (map[c] < 0) map[c] = nCoarse++;
    }
    if (!*forward) {
        for (int c = 0; c < n; c++) map[c] = (nCoarse - 1) - map[c];
    }
    *forward = !*forward;
    *nCoarseOut = nCoarse;
    free(offs); free(cnt); free(cf);
    return map;
}

static void agglomerate_addressing(level_t* fine, level_t* coarse, int nCoarse) {
    const int nF = fine->nFaces;
    const int32_t* rm = fine->restrictAddr;
    fine->faceRestrictAddr = malloc(sizeof(int32_t) * (size_t)(nF > 0 ? nF : 1));
    fine->faceFlip = calloc((size_t)(nF > 0 ? nF : 1), 1);
    int maxN = 10;
    int32_t* cnt = calloc((size_t)nCoarse, sizeof(int32_t));
    int32_t* cfaces = malloc(sizeof(int32_t) * (size_t)maxN * (size_t)nCoarse);
    int32_t* initNei = malloc(sizeof(int32_t) * (size_t)(nF > 0 ? nF : 1));
    int nCF = 0;
    for (int f = 0; f < nF; f++) {
        const int ru = rm[fine->u[f]], rl = rm[fine->l[f]];
        if (ru == rl) { fine->faceRestrictAddr[f] = -(ru + 1); continue; }
        const int own = ru < rl ? ru : rl, nei = ru < rl ? rl : ru;
        int found = -1;
        for (int i = 0; i < cnt[own]; i++) {
            if (initNei[cfaces[(size_t)maxN * own + i]] == nei) { found = cfaces[(size_t)maxN * own + i]; break; }
        }
        if (found < 0) {
            if (cnt[own] >= maxN) {
                const int oldMax = maxN;
                maxN *= 2;
                cfaces = realloc(cfaces, sizeof(int32_t) * (size_t)maxN * (size_t)nCoarse);
                for (int i = nCoarse - 1; i >= 0; i--)
                    for (int j = cnt[i] - 1; j >= 0; j--) cfaces[(size_t)maxN * i + j] = cfaces[(size_t)oldMax * i + j];
            }
            found = nCF++;
            cfaces[(size_t)maxN * own + cnt[own]++] = found;
            initNei[found] = nei;
        }
        fine->faceRestrictAddr[f] = found;
    }
    coarse->nCells = nCoarse;
    coarse->nFaces = nCF;
    coarse->l = malloc(sizeof(int32_t) * (size_t)(nCF > 0 ? nCF : 1));
    coarse->u = malloc(sizeof(int32_t) * (size_t)(nCF > 0 ? nCF : 1));
    int32_t* renum = malloc(sizeof(int32_t) * (size_t)(nCF > 0 ? nCF : 1));
    int k = 0;
    for (int c = 0; c < nCoarse; c++) {
        for (int i = 0; i < cnt[c]; i++) {
            const int e = cfaces[(size_t)maxN * c + i];
            coarse->l[k] = c;
            coarse->u[k] = initNei[e];
            renum[e] = k++;
        }
    }
    for (int f = 0; f < nF; f++) {
        if (fine->faceRestrictAddr[f] >= 0) {
            const int cf = renum[fine->faceRestrictAddr[f]];
            fine->faceRestrictAddr[f] = cf;
            fine->faceFlip[f] = (coarse->l[cf] == rm[fine->u[f]]) ? 1 : 0;
        }
    }
    free(cnt); free(cfaces); free(initNei); free(renum);
}

/* cyclicGAMGInterface constructor (solvers/GAMG/interfaces/cyclicGAMGInterface/cyclicGAMGInterface.C:85-157; the same
 * pair logic as processorGAMGInterface.C:75-147): coarse patch faces are the distinct (master coarse cell, slave coarse
 * cell) pairs in first-seen order; the owner patch (lower patch index) is the master. */
static void agglomerate_interfaces(level_t* fine, level_t* coarse) {
    const int32_t* rm = fine->restrictAddr;
    coarse->nIfaces = fine->nIfaces;
    coarse->ifs = calloc((size_t)(fine->nIfaces > 0 ? fine->nIfaces : 1), sizeof(lev_iface_t));
    coarse->views = calloc((size_t)(fine->nIfaces > 0 ? fine->nIfaces : 1), sizeof(iface_t));
    for (int i = 0; i < fine->nIfaces; i++) {
        lev_iface_t* F = &fine->ifs[i];
        lev_iface_t* Cc = &coarse->ifs[i];
        const int owner = i < F->nbrPatch;
        size_t cap = 16;
        while (cap < (size_t)F->n * 4) cap *= 2;
        int64_t* keys = malloc(sizeof(int64_t) * cap);
        int32_t* vals = malloc(sizeof(int32_t) * cap);
        for (size_t k = 0; k < cap; k++) keys[k] = -1;
        F->faceRestrict = malloc(sizeof(int32_t) * (size_t)(F->n > 0 ? F->n : 1));
        Cc->faceCells = malloc(sizeof(int32_t) * (size_t)(F->n > 0 ? F->n : 1));
        Cc->nbrPatch = F->nbrPatch;
        int nC = 0;
        for (int e = 0; e < F->n; e++) {
            const int32_t loc = rm[F->faceCells[e]], nbr = rm[F->nbrCells[e]];   /* interfaceInternalField / internalFieldTransfer */
            const int64_t key = owner ? (((int64_t)loc << 32) | (uint32_t)nbr) : (((int64_t)nbr << 32) | (uint32_t)loc);
            size_t h = (size_t)((uint64_t)key * 0x9E3779B97F4A7C15ull) & (cap - 1);
            while (keys[h] != -1 && keys[h] != key) h = (h + 1) & (cap - 1);
            if (keys[h] == -1) {
                keys[h] = key;
                vals[h] = nC;
                Cc->faceCells[nC++] = loc;
            }
            F->faceRestrict[e] = vals[h];
        }
        Cc->n = nC;
        free(keys);
        free(vals);
    }
    for (int i = 0; i < coarse->nIfaces; i++) {            /* nbrPatch().faceCells() */
        lev_iface_t* Cc = &coarse->ifs[i];
        const lev_iface_t* N = &coarse->ifs[Cc->nbrPatch];
        if (N->n != Cc->n) abort();                         /* both halves must number their coarse faces alike */
        Cc->nbrCells = malloc(sizeof(int32_t) * (size_t)(Cc->n > 0 ? Cc->n : 1));
        memcpy(Cc->nbrCells, N->faceCells, sizeof(int32_t) * (size_t)Cc->n);
    }
}

static void restrict_field(double* cf, const double* ff, const int32_t* map, int nFine, int nCoarse) {
    for (int i = 0; i < nCoarse; i++) cf[i] = 0;          /* GAMGAgglomerationTemplates.C:84-89 */
    for (int i = 0; i < nFine; i++) cf[map[i]] += ff[i];
}

hierarchy_t* oracle_gamg_build_coupled(int32_t nCells, int32_t nFaces, const int32_t* l, const int32_t* u,
                                       const double* faceWeights, int minCells, int forwardStart, int32_t nIfaces,
                                       const int32_t* ifaceSizes, const int32_t* const* ifaceFaceCells,
                                       const int32_t* ifaceNbrPatch);

hierarchy_t* oracle_gamg_build(int32_t nCells, int32_t nFaces, const int32_t* l, const int32_t* u,
                               const double* faceWeights, int minCells, int forwardStart) {
    return oracle_gamg_build_coupled(nCells, nFaces, l, u, faceWeights, minCells, forwardStart, 0, NULL, NULL, NULL);
}

hierarchy_t* oracle_gamg_build_coupled(int32_t nCells, int32_t nFaces, const int32_t* l, const int32_t* u,
                                       const double* faceWeights, int minCells, int forwardStart, int32_t nIfaces,
                                       const int32_t* ifaceSizes, const int32_t* const* ifaceFaceCells,
                                       const int32_t* ifaceNbrPatch) {
    hierarchy_t* H = calloc(1, sizeof(hierarchy_t));
    level_t* L0 = &H->lev[0];
    L0->nCells = nCells;
    L0->nFaces = nFaces;
    L0->l = malloc(sizeof(int32_t) * (size_t)(nFaces > 0 ? nFaces : 1));
    L0->u = malloc(sizeof(int32_t) * (size_t)(nFaces > 0 ? nFaces : 1));
    memcpy(L0->l, l, sizeof(int32_t) * (size_t)nFaces);
    memcpy(L0->u, u, sizeof(int32_t) * (size_t)nFaces);
    L0->nIfaces = nIfaces;
    L0->ifs = calloc((size_t)(nIfaces > 0 ? nIfaces : 1), sizeof(lev_iface_t));
    L0->views = calloc((size_t)(nIfaces > 0 ? nIfaces : 1), sizeof(iface_t));
    for (int i = 0; i < nIfaces; i++) {
        lev_iface_t* I = &L0->ifs[i];
        const int32_t n = ifaceSizes[i];
        I->n = n;
        I->nbrPatch = ifaceNbrPatch[i];
        I->faceCells = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        I->nbrCells = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        memcpy(I->faceCells, ifaceFaceCells[i], sizeof(int32_t) * (size_t)n);
        memcpy(I->nbrCells, ifaceFaceCells[ifaceNbrPatch[i]], sizeof(int32_t) * (size_t)n);
    }
    H->nLevels = 1;
    int forward = forwardStart;
    double* w = malloc(sizeof(double) * (size_t)(nFaces > 0 ? nFaces : 1));
    memcpy(w, faceWeights, sizeof(double) * (size_t)nFaces);
    while (H->nLevels - 1 < 50 - 1) {                      /* maxLevels_ 50 */
        level_t* fine = &H->lev[H->nLevels - 1];
        int nCoarse;
        int32_t* map = pair_agglomerate(&nCoarse, fine->nCells, fine->nFaces, fine->l, fine->u, w, &forward);
        if (nCoarse < minCells || !(nCoarse < fine->nCells)) {  /* continueAgglomerating, one rank */
            free(map);
            break;
        }
        fine->restrictAddr = map;
        level_t* coarse = &H->lev[H->nLevels];
        agglomerate_addressing(fine, coarse, nCoarse);
        agglomerate_interfaces(fine, coarse);
        double* cw = calloc((size_t)(coarse->nFaces > 0 ? coarse->nFaces : 1), sizeof(double));
        for (int f = 0; f < fine->nFaces; f++)              /* restrictFaceField */
            if (fine->faceRestrictAddr[f] >= 0) cw[fine->faceRestrictAddr[f]] += w[f];
        free(w);
        w = cw;
        H->nLevels++;
    }
    free(w);
    return H;
}

int oracle_gamg_n_levels(const hierarchy_t* H) { return H->nLevels; }
void oracle_gamg_level_sizes(const hierarchy_t* H, int lev, int32_t* nCells, int32_t* nFaces) {
    *nCells = H->lev[lev].nCells;
    *nFaces = H->lev[lev].nFaces;
}
void oracle_gamg_level_arrays(const hierarchy_t* H, int lev, int32_t* restrictAddr, int32_t* faceRestrictAddr,
                              int32_t* faceFlip, int32_t* coarseLower, int32_t* coarseUpper) {
    const level_t* f = &H->lev[lev];
    const level_t* c = &H->lev[lev + 1];
    memcpy(restrictAddr, f->restrictAddr, sizeof(int32_t) * (size_t)f->nCells);
    memcpy(faceRestrictAddr, f->faceRestrictAddr, sizeof(int32_t) * (size_t)f->nFaces);
    for (int i = 0; i < f->nFaces; i++) faceFlip[i] = f->faceFlip[i];
    memcpy(coarseLower, c->l, sizeof(int32_t) * (size_t)c->nFaces);
    memcpy(coarseUpper, c->u, sizeof(int32_t) * (size_t)c->nFaces);
}

/* coarse patch `iface` of level lev+1: its size, faceCells and the fine patch's faceRestrictAddressing */
int oracle_gamg_iface_size(const hierarchy_t* H, int lev, int iface) { return H->lev[lev].ifs[iface].n; }
void oracle_gamg_iface_arrays(const hierarchy_t* H, int lev, int iface, int32_t* coarseFaceCells,
                              int32_t* faceRestrictAddr) {
    const lev_iface_t* F = &H->lev[lev].ifs[iface];
    const lev_iface_t* Cc = &H->lev[lev + 1].ifs[iface];
    memcpy(coarseFaceCells, Cc->faceCells, sizeof(int32_t) * (size_t)Cc->n);
    memcpy(faceRestrictAddr, F->faceRestrict, sizeof(int32_t) * (size_t)F->n);
}

static ldu_t as_ldu(const level_t* L) {
    for (int i = 0; i < L->nIfaces; i++) {
        const lev_iface_t* I = &L->ifs[i];
        iface_t v = {I->n, I->faceCells, I->nbrCells, I->bou, I->inn};
        L->views[i] = v;
    }
    ldu_t A = {L->nCells, L->nFaces, L->l, L->u, L->diag, L->upper, L->lower, L->nIfaces, L->views};
    return A;
}

/* interfaceBouCoeffs / interfaceIntCoeffs of the finest level (copied); call before set_matrix / solve */
void oracle_gamg_set_interface_coeffs(hierarchy_t* H, const double* const* bou, const double* const* inn) {
    level_t* L0 = &H->lev[0];
    for (int i = 0; i < L0->nIfaces; i++) {
        lev_iface_t* I = &L0->ifs[i];
        free(I->bou); free(I->inn);
        I->bou = malloc(sizeof(double) * (size_t)(I->n > 0 ? I->n : 1));
        I->inn = malloc(sizeof(double) * (size_t)(I->n > 0 ? I->n : 1));
        memcpy(I->bou, bou[i], sizeof(double) * (size_t)I->n);
        memcpy(I->inn, inn[i], sizeof(double) * (size_t)I->n);
    }
}

/* agglomerateMatrix for all levels (GAMGSolver.C:196-208, GAMGSolverAgglomerateMatrix.C:33-193) */
static void gamg_set_matrix(hierarchy_t* H, const double* diag, const double* upper, const double* lower) {
    H->symmetric = (lower == NULL);
    for (int k = 0; k < H->nLevels; k++) {
        level_t* L = &H->lev[k];
        if (L->lower && L->lower != L->upper) free(L->lower);
        free(L->diag); free(L->upper); free(L->corr); free(L->src);
        L->diag = malloc(sizeof(double) * (size_t)(L->nCells > 0 ? L->nCells : 1));
        L->upper = calloc((size_t)(L->nFaces > 0 ? L->nFaces : 1), sizeof(double));
        L->lower = H->symmetric ? L->upper : calloc((size_t)(L->nFaces > 0 ? L->nFaces : 1), sizeof(double));
        L->corr = calloc((size_t)(L->nCells > 0 ? L->nCells : 1), sizeof(double));
        L->src = calloc((size_t)(L->nCells > 0 ? L->nCells : 1), sizeof(double));
    }
    level_t* L0 = &H->lev[0];
    memcpy(L0->diag, diag, sizeof(double) * (size_t)L0->nCells);
    memcpy(L0->upper, upper, sizeof(double) * (size_t)L0->nFaces);
    if (!H->symmetric) memcpy(L0->lower, lower, sizeof(double) * (size_t)L0->nFaces);
    for (int k = 0; k + 1 < H->nLevels; k++) {
        level_t *F = &H->lev[k], *C = &H->lev[k + 1];
        for (int i = 0; i < F->nIfaces; i++) {             /* agglomerateInterfaceCoefficients, GAMGSolverAgglomerateMatrix.C:196-275 */
            lev_iface_t *fi = &F->ifs[i], *ci = &C->ifs[i];
            free(ci->bou); free(ci->inn);
            ci->bou = calloc((size_t)(ci->n > 0 ? ci->n : 1), sizeof(double));
            ci->inn = calloc((size_t)(ci->n > 0 ? ci->n : 1), sizeof(double));
            for (int e = 0; e < fi->n; e++) {
                ci->bou[fi->faceRestrict[e]] += fi->bou[e];
                ci->inn[fi->faceRestrict[e]] += fi->inn[e];
            }
        }
        restrict_field(C->diag, F->diag, F->restrictAddr, F->nCells, C->nCells);
        for (int f = 0; f < F->nFaces; f++) {
            const int cf = F->faceRestrictAddr[f];
            if (!H->symmetric) {
                if (cf >= 0) {
                    if (!F->faceFlip[f]) { C->upper[cf] += F->upper[f]; C->lower[cf] += F->lower[f]; }
                    else { C->upper[cf] += F->lower[f]; C->lower[cf] += F->upper[f]; }
                } else {
                    C->diag[-1 - cf] += F->upper[f] + F->lower[f];
                }
            } else {
                if (cf >= 0) C->upper[cf] += F->upper[f];
                else C->diag[-1 - cf] += 2 * F->upper[f];
            }
        }
    }
}

static void gamg_scale(const level_t* L, double* field, double* Acf, const double* source) {
    ldu_t A = as_ldu(L);
    oracle_amul(&A, field, Acf);
    double num = 0, den = 0;
    for (int i = 0; i < L->nCells; i++) {
        num += source[i] * field[i];
        den += Acf[i] * field[i];
    }
    const double sf = num / (den >= 0 ? den + VSMALL : den - VSMALL);
    for (int i = 0; i < L->nCells; i++) field[i] = sf * field[i] + (source[i] - sf * Acf[i]) / L->diag[i];
}

/* one V-cycle (GAMGSolverSolve.C:148-443) */
static void gamg_vcycle(hierarchy_t* H, const ctl_t* c, double* psi, const double* source, double* Apsi,
                        double* finestCorrection, double* finestResidual, double* scratch) {
    const level_t* L0 = &H->lev[0];
    const int n = L0->nCells;
    const int coarsest = H->nLevels - 2;                   /* index into the reference's matrixLevels_ */
    const int scale = c->scaleCorrection < 0 ? H->symmetric : c->scaleCorrection;
    ldu_t A0 = as_ldu(L0);
    ctl_t cc = *c;                                         /* coarsest solver: same tolerance/relTol, defaults else */
    cc.maxIter = 1000;
    cc.minIter = 0;
    cc.precond = H->symmetric ? 2 : 3;
    restrict_field(H->lev[1].src, finestResidual, L0->restrictAddr, n, H->lev[1].nCells);
    for (int l = 0; l < coarsest; l++) {
        level_t* L = &H->lev[l + 1];
        if (c->nPreSweeps) {
            ldu_t A = as_ldu(L);
            for (int i = 0; i < L->nCells; i++) L->corr[i] = 0;
            int ns = c->nPreSweeps + c->preSweepsLevelMultiplier * l;
            if (ns > c->maxPreSweeps) ns = c->maxPreSweeps;
            oracle_smooth(&A, c->precond, L->corr, L->src, ns);
            if (scale && l < coarsest - 1) gamg_scale(L, L->corr, scratch, L->src);
            oracle_amul(&A, L->corr, scratch);
            for (int i = 0; i < L->nCells; i++) L->src[i] -= scratch[i];
        }
        restrict_field(H->lev[l + 2].src, L->src, L->restrictAddr, L->nCells, H->lev[l + 2].nCells);
    }
    {   /* solveCoarsestLevel */
        level_t* L = &H->lev[coarsest + 1];
        ldu_t A = as_ldu(L);
        perf_t cp;
        for (int i = 0; i < L->nCells; i++) L->corr[i] = 0;
        if (L->nFaces == 0) {
            for (int i = 0; i < L->nCells; i++) L->corr[i] = L->src[i] / L->diag[i];
        } else if (H->symmetric) {
            oracle_pcg(&A, &cc, L->corr, L->src, &cp);
        } else {
            oracle_pbicgstab(&A, &cc, L->corr, L->src, &cp);
        }
    }
    for (int l = coarsest - 1; l >= 0; l--) {
        level_t* L = &H->lev[l + 1];
        ldu_t A = as_ldu(L);
        double* pre = NULL;
        if (c->nPreSweeps) {
            pre = malloc(sizeof(double) * (size_t)L->nCells);
            memcpy(pre, L->corr, sizeof(double) * (size_t)L->nCells);
        }
        for (int i = 0; i < L->nCells; i++) L->corr[i] = H->lev[l + 2].corr[L->restrictAddr[i]];
        if (scale && l < coarsest - 1) gamg_scale(L, L->corr, scratch, L->src);
        if (pre) {
            for (int i = 0; i < L->nCells; i++) L->corr[i] += pre[i];
            free(pre);
        }
        int ns = c->nPostSweeps + c->postSweepsLevelMultiplier * l;
        if (ns > c->maxPostSweeps) ns = c->maxPostSweeps;
        oracle_smooth(&A, c->precond, L->corr, L->src, ns);
    }
    for (int i = 0; i < n; i++) finestCorrection[i] = H->lev[1].corr[L0->restrictAddr[i]];
    if (scale) gamg_scale(L0, finestCorrection, Apsi, finestResidual);
    for (int i = 0; i < n; i++) psi[i] += finestCorrection[i];
    oracle_smooth(&A0, c->precond, psi, source, c->nFinestSweeps);
}

void oracle_gamg_set_matrix(hierarchy_t* H, const double* diag, const double* upper, const double* lower) {
    gamg_set_matrix(H, diag, upper, lower);
}

/* GAMGPreconditioner::precondition (preconditioners/GAMGPreconditioner/GAMGPreconditioner.C:81-148) */
static void gamg_precondition(const ctl_t* c, const double* rA, double* wA) {
    hierarchy_t* H = (hierarchy_t*)c->hierarchy;
    const int n = H->lev[0].nCells;
    ldu_t A0 = as_ldu(&H->lev[0]);
    ctl_t g = *c;
    g.precond = c->precSmoother;
    g.tolerance = c->precTolerance;
    g.relTol = c->precRelTol;
    double* AwA = malloc(sizeof(double) * (size_t)n * 4);
    double *finestCorrection = AwA + n, *finestResidual = finestCorrection + n, *scratch = finestResidual + n;
    for (int i = 0; i < n; i++) { wA[i] = 0.0; finestResidual[i] = rA[i]; }
    for (int cycle = 0; cycle < g.nVcycles; cycle++) {
        gamg_vcycle(H, &g, wA, rA, AwA, finestCorrection, finestResidual, scratch);
        if (cycle < g.nVcycles - 1) {
            oracle_amul(&A0, wA, AwA);
            for (int i = 0; i < n; i++) finestResidual[i] = rA[i] - AwA[i];
        }
    }
    free(AwA);
}

/* GAMGSolver::solve (GAMGSolverSolve.C:31-145) */
void oracle_gamg_solve(hierarchy_t* H, const double* diag, const double* upper, const double* lower,
                       const ctl_t* c, double* psi, const double* source, perf_t* perf) {
    gamg_set_matrix(H, diag, upper, lower);
    const level_t* L0 = &H->lev[0];
    const int n = L0->nCells;
    ldu_t A0 = as_ldu(L0);
    double* Apsi = malloc(sizeof(double) * (size_t)n * 4);
    double *finestCorrection = Apsi + n, *finestResidual = finestCorrection + n, *scratch = finestResidual + n;
    memset(perf, 0, sizeof(*perf));
    oracle_amul(&A0, psi, Apsi);
    const double nf = oracle_norm_factor(&A0, psi, source, Apsi, finestCorrection);
    perf->normFactor = nf;
    for (int i = 0; i < n; i++) finestResidual[i] = source[i] - Apsi[i];
    perf->initialResidual = sum_mag(finestResidual, n) / nf;
    perf->finalResidual = perf->initialResidual;
    if (c->minIter > 0 || !check(perf, c->tolerance, c->relTol)) {
        do {
            gamg_vcycle(H, c, psi, source, Apsi, finestCorrection, finestResidual, scratch);
            oracle_amul(&A0, psi, Apsi);
            for (int i = 0; i < n; i++) finestResidual[i] = source[i] - Apsi[i];
            perf->finalResidual = sum_mag(finestResidual, n) / nf;
            record(perf);
        } while ((++perf->nIterations < c->maxIter && !check(perf, c->tolerance, c->relTol)) ||
                 perf->nIterations < c->minIter);
    }
    free(Apsi);
}

void oracle_gamg_free(hierarchy_t* H) {
    for (int k = 0; k < H->nLevels; k++) {
        level_t* L = &H->lev[k];
        if (L->lower && L->lower != L->upper) free(L->lower);
        free(L->l); free(L->u); free(L->diag); free(L->upper); free(L->restrictAddr); free(L->faceRestrictAddr);
        free(L->faceFlip); free(L->corr); free(L->src);
        for (int i = 0; i < L->nIfaces; i++) {
            lev_iface_t* I = &L->ifs[i];
            free(I->faceCells); free(I->nbrCells); free(I->faceRestrict); free(I->bou); free(I->inn);
        }
        free(L->ifs); free(L->views);
    }
    free(H);
}
