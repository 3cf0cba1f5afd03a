/*---------------------------------------------------------------------------*\
  libB200LinearSolvers -- OpenFOAM-dev plugin: lduMatrix solvers running on a
  B200 through libb200ls (include/b200ls.h).

  Selected with only the fvSolution `solver` keyword and a controlDict `libs`
  entry:

      // system/controlDict
      libs ("libB200LinearSolvers.so");

      // system/fvSolution
      p { solver B200PCG;  preconditioner DIC;  tolerance 1e-6; relTol 0.01; }
      U { solver B200PBiCGStab; preconditioner DILU; tolerance 1e-5; relTol 0.1; }
      p { solver B200GAMG; smoother GaussSeidel; tolerance 1e-6; relTol 0.01; }

  Registration follows the reference's own pattern (PCG.C:34-35,
  PBiCGStab.C:34-38, GAMGSolver.C:38-42): namespace-scope
  lduMatrix::solver::add{sym,asym}MatrixConstructorToTable<T> objects insert
  T::typeName -> T::New into the run-time selection tables when the library is
  dlopen'ed by Time (db/Time/Time.C:403 -> dlLibraryTable::open).

  This file holds no arithmetic: it marshals the lduMatrix the solver was
  constructed with (lduMatrix.H:187-195) into the C-ABI and the returned
  b200ls_perf into a solverPerformance.  Everything persistent lives in a
  process-level cache keyed on the lduAddressing object, because OpenFOAM
  constructs a new solver for every solve (fvScalarMatrix.C:162-170).
\*---------------------------------------------------------------------------*/

#include "lduMatrix.H"
#include "processorLduInterface.H"
#include "cyclicLduInterface.H"
#include "cyclicLduInterfaceField.H"
#include "processorLduInterfaceField.H"
#include "GAMGAgglomeration.H"
#include "addToRunTimeSelectionTable.H"
#include "Pstream.H"
#include "OSspecific.H"

#include "b200ls.h"

#include <map>
#include <string>
#include <cstdio>
#include "OStringStream.H"
#include <vector>
#include <cstdlib>
#include <cstdint>

namespace Foam
{

// ---------------------------------------------------------------------------
// process-level state
// ---------------------------------------------------------------------------

namespace B200
{

struct cacheEntry
{
    b200ls_mesh_t mesh;
    b200ls_matrix_t matrix;
    label nCells;
    label nFaces;
    uint64_t fingerprint;
    //- The reference's agglomeration object handed to the library
    //  (nullptr: none yet).  A rebuilt GAMGAgglomeration (e.g.
    //  cacheAgglomeration false) is a new object and is handed over again
    const void* agglomeration;
    //- Serial of the preconditioner/smoother/solver object whose
    //  coefficients the device matrix holds
    uint64_t owner;
    //- Last use (for eviction)
    uint64_t stamp;
    //- Live preconditioner/smoother objects holding a reference
    int pins;

    cacheEntry()
    :
        mesh(nullptr),
        matrix(nullptr),
        nCells(-1),
        nFaces(-1),
        fingerprint(0),
        agglomeration(nullptr),
        owner(0),
        stamp(0),
        pins(0)
    {}
};

static std::map<const lduAddressing*, cacheEntry> cache_;
static bool initialised_ = false;
static uint64_t clock_ = 0;
static uint64_t nextOwner_ = 0;

//- At most this many addressings keep device meshes alive (dynamic or
//  topology-changing meshes create new lduAddressing objects)
static const size_t maxCacheEntries_ = 16;

static void freeEntry(cacheEntry& e)
{
    if (e.matrix) b200ls_matrix_free(e.matrix);
    if (e.mesh) b200ls_mesh_free(e.mesh);
    e = cacheEntry();
}


//- Fingerprint of the addressing (sizes + every lower/upper label; a few
//  ms per million faces, small against a solve): a mesh that changed
//  topology under an unchanged lduAddressing address must not reuse the
//  cached device layout
static inline uint64_t fpMix(uint64_t h, const uint64_t v)
{
    h ^= v*0x9E3779B97F4A7C15ull;
    h = (h << 27) | (h >> 37);
    return h*0x94D049BB133111EBull + 0x2545F4914F6CDD1Dull;
}

static uint64_t fingerprint(const labelUList& l, uint64_t seed)
{
    // four independent lanes over pairs of labels (memory-bound)
    uint64_t h0 = seed, h1 = ~seed, h2 = seed ^ 0xA5A5A5A5A5A5A5A5ull, h3 = seed + 0x632BE59BD9B4E019ull;
    const label n = l.size();
    label i = 0;
    for (; i + 8 <= n; i += 8)
    {
        h0 = fpMix(h0, (uint64_t(uint32_t(l[i])) << 32) | uint32_t(l[i+1]));
        h1 = fpMix(h1, (uint64_t(uint32_t(l[i+2])) << 32) | uint32_t(l[i+3]));
        h2 = fpMix(h2, (uint64_t(uint32_t(l[i+4])) << 32) | uint32_t(l[i+5]));
        h3 = fpMix(h3, (uint64_t(uint32_t(l[i+6])) << 32) | uint32_t(l[i+7]));
    }
    for (; i < n; i++) h0 = fpMix(h0, uint32_t(l[i]));
    return fpMix(fpMix(fpMix(fpMix(uint64_t(n), h0), h1), h2), h3);
}

static uint64_t fingerprint(const lduAddressing& addr)
{
    return fpMix
    (
        fingerprint(addr.lowerAddr(), 1) ^ uint64_t(addr.size()),
        fingerprint(addr.upperAddr(), 2)
    );
}


static void check(const int rc, const char* what)
{
    if (rc != 0)
    {
        FatalErrorInFunction
            << what << " failed: " << b200ls_last_error()
            << exit(FatalError);
    }
}


//- One process per GPU: device = B200LS_DEVICE if set, else the node-local
//  rank the MPI launcher exports (OpenMPI, Slurm, PMI/MPICH/Hydra, MVAPICH,
//  Intel MPI) modulo the number of visible devices.  Two ranks of one host
//  on the same device are an error (the persistent sweep kernels and the
//  peer-memory reductions of several processes would time-slice on one GPU).
//  In parallel the NCCL id is created on the master and scattered through
//  Pstream.
static void init()
{
    if (initialised_) return;

    int device = 0;
    if (const char* s = getenv("B200LS_DEVICE"))
    {
        device = atoi(s);
    }
    else
    {
        static const char* const vars[] =
        {
            "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", "PMI_LOCAL_RANK",
            "MPI_LOCALRANKID", "MV2_COMM_WORLD_LOCAL_RANK"
        };
        for (const char* v : vars)
        {
            if (const char* s = getenv(v))
            {
                device = atoi(s);
                break;
            }
        }
        const int nDev = b200ls_device_count();
        if (nDev > 0) device %= nDev;
    }

    if (Pstream::parRun())
    {
        // (host, device) of every rank: no two ranks may share a GPU
        List<string> where(Pstream::nProcs());
        where[Pstream::myProcNo()] = hostName() + ":" + Foam::name(device);
        Pstream::gatherList(where);
        Pstream::scatterList(where);
        forAll(where, proci)
        {
            if (proci != Pstream::myProcNo() && where[proci] == where[Pstream::myProcNo()])
            {
                FatalErrorInFunction
                    << "ranks " << proci << " and " << Pstream::myProcNo()
                    << " both map to GPU " << where[proci]
                    << ": libB200LinearSolvers runs one rank per GPU (set"
                    << " B200LS_DEVICE per rank, or start at most as many"
                    << " ranks per host as it has GPUs)" << exit(FatalError);
            }
        }
    }

    if (Pstream::parRun())
    {
        List<char> id(128, '\0');
        if (Pstream::master())
        {
            check(b200ls_nccl_unique_id(id.begin()), "b200ls_nccl_unique_id");
        }
        Pstream::scatter(id);
        check
        (
            b200ls_init(device, id.begin(), Pstream::myProcNo(), Pstream::nProcs()),
            "b200ls_init"
        );
    }
    else
    {
        check(b200ls_init(device, nullptr, 0, 1), "b200ls_init");
    }

    initialised_ = true;
}


static int preconditionerId(const word& name)
{
    if (name == "DIC") return B200LS_DIC;
    if (name == "DILU") return B200LS_DILU;
    if (name == "diagonal") return B200LS_DIAGONAL;
    if (name == "none") return B200LS_NONE;
    if (name == "GaussSeidel") return B200LS_GAUSS_SEIDEL;
    if (name == "symGaussSeidel") return B200LS_SYM_GAUSS_SEIDEL;
    if (name == "DICGaussSeidel") return B200LS_DIC_GAUSS_SEIDEL;
    if (name == "DILUGaussSeidel") return B200LS_DILU_GAUSS_SEIDEL;
    if (name == "GAMG") return B200LS_GAMG_PRECOND;

    FatalErrorInFunction
        << "preconditioner/smoother " << name
        << " is not provided by libB200LinearSolvers."
        << " Valid: DIC DILU diagonal none GAMG (preconditioners),"
        << " GaussSeidel symGaussSeidel DIC DILU DICGaussSeidel"
        << " DILUGaussSeidel (smoothers)"
        << exit(FatalError);
    return -1;
}

//- Mesh/matrix handles for an addressing (created on demand, cached)
static cacheEntry& entryFor
(
    const lduMatrix& matrix,
    const lduInterfaceFieldPtrsList& interfaces
)
{
    init();

    const lduAddressing& addr = matrix.lduAddr();
    if (cache_.find(&addr) == cache_.end() && cache_.size() >= maxCacheEntries_)
    {
        // evict the entry that has not been used for the longest time
        auto oldest = cache_.end();
        for (auto it = cache_.begin(); it != cache_.end(); ++it)
        {
            if
            (
                it->second.pins == 0
             && (oldest == cache_.end() || it->second.stamp < oldest->second.stamp)
            )
            {
                oldest = it;
            }
        }
        if (oldest != cache_.end())
        {
            freeEntry(oldest->second);
            cache_.erase(oldest);
        }
    }
    cacheEntry& e = cache_[&addr];
    e.stamp = ++clock_;

    const label nCells = addr.size();
    const label nFaces = addr.lowerAddr().size();

    // (hashing every label on every call would cost as much as a small
    //  solve: sizes are compared always, the labels when the entry is created
    //  and whenever the sizes of a cached entry still match)
    const uint64_t fp = fingerprint(addr);

    if (e.mesh && (e.nCells != nCells || e.nFaces != nFaces || e.fingerprint != fp))
    {
        // mesh changed under the same address: rebuild
        const int pins = e.pins;
        freeEntry(e);
        e.stamp = clock_;
        e.pins = pins;
    }

    if (!e.mesh)
    {
        std::vector<int32_t> sizes, nbr;
        std::vector<const int32_t*> faceCells;

        // coupled patches in patch order; cyclic halves refer to their
        // partner by its index among the coupled patches
        std::vector<int32_t> coupledIndex(interfaces.size(), -1);
        forAll(interfaces, patchi)
        {
            if (interfaces.set(patchi))
            {
                coupledIndex[patchi] = sizes.size();
                sizes.push_back(0);
            }
        }
        sizes.clear();

        forAll(interfaces, patchi)
        {
            if (interfaces.set(patchi))
            {
                const lduInterface& li = interfaces[patchi].interface();
                const labelUList& fc = li.faceCells();
                sizes.push_back(fc.size());
                faceCells.push_back(fc.begin());

                if (isA<processorLduInterface>(li))
                {
                    if
                    (
                        isA<processorLduInterfaceField>(interfaces[patchi])
                     && refCast<const processorLduInterfaceField>
                        (
                            interfaces[patchi]
                        ).transforms()
                    )
                    {
                        FatalErrorInFunction
                            << "processor patch " << patchi
                            << " transforms the field (processorCyclic half"
                            << " of a rotational cyclic on a vector/tensor"
                            << " component): not supported by"
                            << " libB200LinearSolvers" << exit(FatalError);
                    }
                    nbr.push_back
                    (
                        refCast<const processorLduInterface>(li)
                       .neighbProcNo()
                    );
                }
                else if (isA<cyclicLduInterface>(li))
                {
                    if
                    (
                        isA<cyclicLduInterfaceField>(interfaces[patchi])
                     && refCast<const cyclicLduInterfaceField>
                        (
                            interfaces[patchi]
                        ).transforms()
                    )
                    {
                        FatalErrorInFunction
                            << "cyclic patch " << patchi
                            << " transforms the field (rotational cyclic"
                            << " on a vector/tensor component): not"
                            << " supported by libB200LinearSolvers"
                            << exit(FatalError);
                    }
                    const label nbrPatch =
                        refCast<const cyclicLduInterface>(li)
                       .nbrPatchIndex();
                    nbr.push_back(B200LS_CYCLIC(coupledIndex[nbrPatch]));
                }
                else
                {
                    FatalErrorInFunction
                        << "coupled patch " << patchi << " of type "
                        << li.type() << " is neither a processor nor a"
                        << " cyclic interface"
                        << exit(FatalError);
                }
            }
        }

        e.mesh = b200ls_mesh_create
        (
            nCells,
            nFaces,
            addr.lowerAddr().begin(),
            addr.upperAddr().begin(),
            sizes.size(),
            sizes.data(),
            faceCells.data(),
            nbr.data()
        );
        if (!e.mesh)
        {
            FatalErrorInFunction
                << "b200ls_mesh_create failed: " << b200ls_last_error()
                << exit(FatalError);
        }
        e.matrix = b200ls_matrix_create(e.mesh);
        e.nCells = nCells;
        e.nFaces = nFaces;
        e.fingerprint = fp;
    }

    return e;
}


//- Upload the coefficients a solver/smoother/preconditioner object was
//  constructed with
static void uploadCoeffs
(
    cacheEntry& e,
    const lduMatrix& matrix,
    const Field<Field<scalar>>& bouCoeffs,
    const Field<Field<scalar>>& intCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const uint64_t owner = 0
)
{
    e.owner = owner;

    std::vector<const double*> bou, inn;
    forAll(interfaces, patchi)
    {
        if (interfaces.set(patchi))
        {
            bou.push_back(bouCoeffs[patchi].begin());
            inn.push_back(intCoeffs[patchi].begin());
        }
    }

    // PISO solves the same pressure matrix once per corrector: unchanged
    // coefficients are not uploaded again and the factorisation / GAMG
    // coarse-level matrices on the device stay valid
    int32_t changed = 1;
    check
    (
        b200ls_matrix_set_if_changed
        (
            e.matrix,
            matrix.diag().begin(),
            matrix.upper().begin(),
            matrix.asymmetric() ? matrix.lower().begin() : nullptr,
            bou.data(),
            inn.data(),
            &changed
        ),
        "b200ls_matrix_set_if_changed"
    );
}

} // End namespace B200


// ---------------------------------------------------------------------------
// common base: marshalling
// ---------------------------------------------------------------------------

class B200SolverBase
:
    public lduMatrix::solver
{
protected:

    //- Mesh/matrix handles for this solver's addressing (created on demand)
    B200::cacheEntry& entry() const
    {
        return B200::entryFor(matrix_, interfaces_);
    }

    //- Upload the coefficients this solver object was constructed with
    void upload(B200::cacheEntry& e) const
    {
        B200::uploadCoeffs
        (
            e, matrix_, interfaceBouCoeffs_, interfaceIntCoeffs_, interfaces_
        );
    }

    //- Hand the reference's cached agglomeration (GAMGAgglomeration::New,
    //  GAMGAgglomeration.C:349-400) to the library, once per mesh
    void ensureAgglomeration(B200::cacheEntry& e, const dictionary& dict) const
    {
        if (dict.found("processorAgglomerator"))
        {
            FatalErrorInFunction
                << "processorAgglomerator "
                << word(dict.lookup("processorAgglomerator"))
                << ": processor agglomeration is not supported by"
                << " libB200LinearSolvers" << exit(FatalError);
        }

        // the reference's (cached) agglomeration; a rebuilt one is a new
        // object and is handed to the library again
        const GAMGAgglomeration& agg = GAMGAgglomeration::New(matrix_, dict);
        if (e.agglomeration == &agg) return;

        std::vector<const int32_t*> maps;
        std::vector<int32_t> nCoarse;
        label nFine = matrix_.lduAddr().size();
        for (label lev = 0; lev < agg.size(); lev++)
        {
            // restrictAddressing(lev): one coarse label per cell of level lev
            if (agg.restrictAddressing(lev).size() != nFine)
            {
                FatalErrorInFunction
                    << "restrictAddressing of level " << lev << " has "
                    << agg.restrictAddressing(lev).size() << " entries for "
                    << nFine << " cells" << exit(FatalError);
            }
            maps.push_back(agg.restrictAddressing(lev).begin());
            nCoarse.push_back(agg.nCells(lev));
            nFine = agg.nCells(lev);
        }
        if
        (
            b200ls_agglomerate_from_maps
            (
                e.mesh,
                maps.size(),
                maps.data(),
                nCoarse.data()
            ) < 0
        )
        {
            FatalErrorInFunction
                << "b200ls_agglomerate_from_maps failed: "
                << b200ls_last_error() << exit(FatalError);
        }
        e.agglomeration = &agg;
    }

    //- GAMGSolver::readControls (GAMGSolver.C:348-371)
    void readGAMGControls(b200ls_controls& c, const dictionary& dict) const
    {
        c.nPreSweeps = dict.lookupOrDefault<label>("nPreSweeps", 0);
        c.preSweepsLevelMultiplier =
            dict.lookupOrDefault<label>("preSweepsLevelMultiplier", 1);
        c.maxPreSweeps = dict.lookupOrDefault<label>("maxPreSweeps", 4);
        c.nPostSweeps = dict.lookupOrDefault<label>("nPostSweeps", 2);
        c.postSweepsLevelMultiplier =
            dict.lookupOrDefault<label>("postSweepsLevelMultiplier", 1);
        c.maxPostSweeps = dict.lookupOrDefault<label>("maxPostSweeps", 4);
        c.nFinestSweeps = dict.lookupOrDefault<label>("nFinestSweeps", 2);
        c.scaleCorrection =
            dict.lookupOrDefault<Switch>("scaleCorrection", matrix_.symmetric());

        if (dict.lookupOrDefault<Switch>("interpolateCorrection", false))
        {
            FatalErrorInFunction
                << "interpolateCorrection is not supported by libB200LinearSolvers"
                << exit(FatalError);
        }
        if (dict.lookupOrDefault<Switch>("directSolveCoarsest", false))
        {
            FatalErrorInFunction
                << "directSolveCoarsest is not supported by libB200LinearSolvers"
                << exit(FatalError);
        }
    }

    //- Preconditioner entry of a Krylov solver: a word, or a sub-dictionary
    //  (lduMatrixPreconditioner.C:40-59); preconditioner GAMG takes its V-cycle
    //  controls from the sub-dictionary (GAMGPreconditioner.C:47-79)
    word readPreconditioner(b200ls_controls& c) const
    {
        const word precon(lduMatrix::preconditioner::getName(controlDict_));
        c.precond = B200::preconditionerId(precon);

        if (c.precond == B200LS_GAMG_PRECOND)
        {
            const Foam::entry& pe =
                controlDict_.lookupEntry("preconditioner", false, false);
            // a word entry carries no controls: the reference constructs the
            // preconditioner from dictionary::null then
            // (lduMatrixPreconditioner.C:40-59), i.e. every default
            const dictionary& pd = pe.isDict() ? pe.dict() : dictionary::null;

            ensureAgglomeration(entry(), pd);
            readGAMGControls(c, pd);
            c.precSmoother =
                B200::preconditionerId(word(pd.lookup("smoother")));
            c.nVcycles = pd.lookupOrDefault<label>("nVcycles", 2);
            c.precTolerance = pd.lookupOrDefault<scalar>("tolerance", 1e-6);
            c.precRelTol = pd.lookupOrDefault<scalar>("relTol", 0);
        }

        return precon;
    }

    //- Run the solve and translate the result
    solverPerformance run
    (
        const word& solverName,
        b200ls_controls& c,
        scalarField& psi,
        const scalarField& source
    ) const
    {
        B200::cacheEntry& e = entry();
        upload(e);

        c.tolerance = tolerance_;
        c.relTol = relTol_;
        c.maxIter = maxIter_;
        c.minIter = minIter_;

        b200ls_perf* perf = new b200ls_perf;
        B200::check
        (
            b200ls_solve(e.matrix, &c, psi.begin(), source.begin(), perf),
            "b200ls_solve"
        );

        // converged_ is the flag of the last checkConvergence call the
        // reference's loop would have made (SolverPerformance.C:60-92),
        // which libb200ls tracks; do not re-evaluate it here
        solverPerformance solverPerf
        (
            solverName,
            fieldName_,
            perf->initialResidual,
            perf->finalResidual,
            perf->nIterations,
            perf->converged != 0,
            perf->singular != 0
        );
        delete perf;

        return solverPerf;
    }


public:

    B200SolverBase
    (
        const word& fieldName,
        const lduMatrix& matrix,
        const Field<Field<scalar>>& interfaceBouCoeffs,
        const Field<Field<scalar>>& interfaceIntCoeffs,
        const lduInterfaceFieldPtrsList& interfaces,
        const dictionary& solverControls
    )
    :
        lduMatrix::solver
        (
            fieldName,
            matrix,
            interfaceBouCoeffs,
            interfaceIntCoeffs,
            interfaces,
            solverControls
        )
    {}

    virtual ~B200SolverBase()
    {}
};


// ---------------------------------------------------------------------------
// B200PCG  (drop-in for PCG, solvers/PCG/PCG.C)
// ---------------------------------------------------------------------------

class B200PCG
:
    public B200SolverBase
{
public:

    TypeName("B200PCG");

    using B200SolverBase::B200SolverBase;

    virtual solverPerformance solve
    (
        scalarField& psi,
        const scalarField& source,
        const direction cmpt = 0
    ) const
    {
        b200ls_controls c;
        b200ls_controls_default(&c);
        c.solver = B200LS_PCG;
        const word precon(readPreconditioner(c));
        return run(precon + "PCG", c, psi, source);
    }
};


// ---------------------------------------------------------------------------
// B200PBiCGStab  (drop-in for PBiCGStab, solvers/PBiCGStab/PBiCGStab.C)
// ---------------------------------------------------------------------------

class B200PBiCGStab
:
    public B200SolverBase
{
public:

    TypeName("B200PBiCGStab");

    using B200SolverBase::B200SolverBase;

    virtual solverPerformance solve
    (
        scalarField& psi,
        const scalarField& source,
        const direction cmpt = 0
    ) const
    {
        b200ls_controls c;
        b200ls_controls_default(&c);
        c.solver = B200LS_PBICGSTAB;
        const word precon(readPreconditioner(c));
        return run(precon + "PBiCGStab", c, psi, source);
    }
};


// ---------------------------------------------------------------------------
// B200smoothSolver  (drop-in for smoothSolver)
// ---------------------------------------------------------------------------

class B200smoothSolver
:
    public B200SolverBase
{
public:

    TypeName("B200smoothSolver");

    using B200SolverBase::B200SolverBase;

    virtual solverPerformance solve
    (
        scalarField& psi,
        const scalarField& source,
        const direction cmpt = 0
    ) const
    {
        b200ls_controls c;
        b200ls_controls_default(&c);
        c.solver = B200LS_SMOOTH_SOLVER;
        c.precond = B200::preconditionerId(word(controlDict_.lookup("smoother")));
        c.nSweeps = controlDict_.lookupOrDefault<label>("nSweeps", 1);
        return run("smoothSolver", c, psi, source);
    }
};


// ---------------------------------------------------------------------------
// B200GAMG  (drop-in for GAMG, solvers/GAMG/GAMGSolver*.C)
//
// The agglomeration is the reference's own cached MeshObject
// (GAMGAgglomeration::New, GAMGAgglomeration.C:349-400: faceAreaPair by default);
// only its restrictAddressing per level crosses the C-ABI, everything derived is
// rebuilt (and checked bit-exact against the reference) inside libb200ls.
// ---------------------------------------------------------------------------

class B200GAMG
:
    public B200SolverBase
{
public:

    TypeName("B200GAMG");

    using B200SolverBase::B200SolverBase;

    virtual solverPerformance solve
    (
        scalarField& psi,
        const scalarField& source,
        const direction cmpt = 0
    ) const
    {
        ensureAgglomeration(entry(), controlDict_);

        b200ls_controls c;
        b200ls_controls_default(&c);
        c.solver = B200LS_GAMG;
        c.precond = B200::preconditionerId(word(controlDict_.lookup("smoother")));
        readGAMGControls(c, controlDict_);

        return run("GAMG", c, psi, source);
    }
};


// ---------------------------------------------------------------------------
// B200DIC / B200DILU preconditioners and B200* smoothers: the operator-level
// drop-ins (lduMatrix::preconditioner, lduMatrix.H:410-509; lduMatrix::smoother,
// :270-406) for callers that keep the REFERENCE's solver loop, e.g.
//     p { solver PCG;  preconditioner B200DIC; }
//     p { solver GAMG; smoother B200GaussSeidel; }
// Every application crosses the host<->device boundary (wA/rA or psi/source), so
// these are for validation and mixed use; the B200 solvers above keep the whole
// loop on the device.
// ---------------------------------------------------------------------------

template<int Kind>
class B200Preconditioner
:
    public lduMatrix::preconditioner
{
    B200::cacheEntry& e_;
    const uint64_t owner_;

    //- The device matrix of an addressing is shared: if another object has
    //  put its coefficients there since, bring ours back
    void upload() const
    {
        B200::uploadCoeffs
        (
            e_,
            solver_.matrix(),
            solver_.interfaceBouCoeffs(),
            solver_.interfaceIntCoeffs(),
            solver_.interfaces(),
            owner_
        );
    }

public:

    B200Preconditioner(const lduMatrix::solver& sol, const dictionary&)
    :
        lduMatrix::preconditioner(sol),
        e_(B200::entryFor(sol.matrix(), sol.interfaces())),
        owner_(++B200::nextOwner_)
    {
        // coefficients are read once per construction, like
        // DICPreconditioner's calcReciprocalD (DICPreconditioner.C:42-52)
        e_.pins++;
        upload();
    }

    virtual ~B200Preconditioner()
    {
        e_.pins--;
    }

    virtual void precondition
    (
        scalarField& wA,
        const scalarField& rA,
        const direction cmpt = 0
    ) const
    {
        if (e_.owner != owner_) upload();
        B200::check
        (
            b200ls_precondition(e_.matrix, Kind, rA.begin(), wA.begin()),
            "b200ls_precondition"
        );
    }
};

class B200DICPreconditioner
:
    public B200Preconditioner<B200LS_DIC>
{
public:
    TypeName("B200DIC");
    using B200Preconditioner<B200LS_DIC>::B200Preconditioner;
};

class B200DILUPreconditioner
:
    public B200Preconditioner<B200LS_DILU>
{
public:
    TypeName("B200DILU");
    using B200Preconditioner<B200LS_DILU>::B200Preconditioner;
};


template<int Kind>
class B200Smoother
:
    public lduMatrix::smoother
{
    B200::cacheEntry& e_;
    const uint64_t owner_;

    void upload() const
    {
        B200::uploadCoeffs
        (
            e_,
            matrix_,
            interfaceBouCoeffs_,
            interfaceIntCoeffs_,
            interfaces_,
            owner_
        );
    }

public:

    B200Smoother
    (
        const word& fieldName,
        const lduMatrix& matrix,
        const Field<Field<scalar>>& interfaceBouCoeffs,
        const Field<Field<scalar>>& interfaceIntCoeffs,
        const lduInterfaceFieldPtrsList& interfaces
    )
    :
        lduMatrix::smoother
        (
            fieldName,
            matrix,
            interfaceBouCoeffs,
            interfaceIntCoeffs,
            interfaces
        ),
        e_(B200::entryFor(matrix, interfaces)),
        owner_(++B200::nextOwner_)
    {
        e_.pins++;
        upload();
    }

    virtual ~B200Smoother()
    {
        e_.pins--;
    }

    virtual void smooth
    (
        scalarField& psi,
        const scalarField& source,
        const direction cmpt,
        const label nSweeps
    ) const
    {
        if (e_.owner != owner_) upload();
        B200::check
        (
            b200ls_smooth(e_.matrix, Kind, psi.begin(), source.begin(), nSweeps),
            "b200ls_smooth"
        );
    }
};

#define B200_SMOOTHER(ClassName, Kind, Name)                                   \
    class ClassName                                                            \
    :                                                                          \
        public B200Smoother<Kind>                                              \
    {                                                                          \
    public:                                                                    \
        TypeName(Name);                                                        \
        using B200Smoother<Kind>::B200Smoother;                                \
    }

B200_SMOOTHER(B200GaussSeidelSmoother, B200LS_GAUSS_SEIDEL, "B200GaussSeidel");
B200_SMOOTHER
(
    B200symGaussSeidelSmoother, B200LS_SYM_GAUSS_SEIDEL, "B200symGaussSeidel"
);
B200_SMOOTHER(B200DICSmoother, B200LS_DIC, "B200DIC");
B200_SMOOTHER(B200DILUSmoother, B200LS_DILU, "B200DILU");
B200_SMOOTHER
(
    B200DICGaussSeidelSmoother, B200LS_DIC_GAUSS_SEIDEL, "B200DICGaussSeidel"
);
B200_SMOOTHER
(
    B200DILUGaussSeidelSmoother, B200LS_DILU_GAUSS_SEIDEL, "B200DILUGaussSeidel"
);


// ---------------------------------------------------------------------------
// B200dump: pass-through solver that serialises the system it receives at the
// drop-in boundary (SURVEY.md 8(c) "capturing real matrices at the boundary")
// and then delegates to the reference solver named by `delegate`:
//     p { solver B200dump; delegate PCG; dumpFile "p"; preconditioner DIC; ... }
// writes <dumpFile>.<n>.b2ls (addressing, coefficients, coupled patches, psi0,
// source; format: openfoam-dev_b200/ldu_io.py), which replays through
// oracle/ref_harness, the C oracle and the GPU path alike.  Uses no GPU.
// ---------------------------------------------------------------------------

class B200dump
:
    public lduMatrix::solver
{
    struct writer
    {
        FILE* f_;
        std::vector<std::pair<std::string, std::vector<char>>> entries_;
        std::vector<std::pair<int64_t, int64_t>> types_;

        void put(const std::string& name, int64_t dtype, int64_t count, const void* data, size_t bytes)
        {
            entries_.push_back
            (
                std::make_pair
                (
                    name,
                    std::vector<char>
                    (
                        static_cast<const char*>(data),
                        static_cast<const char*>(data) + bytes
                    )
                )
            );
            types_.push_back(std::make_pair(dtype, count));
        }
        void putI32(const std::string& name, const labelUList& l)
        {
            std::vector<int32_t> v(l.size());
            forAll(l, i) v[i] = l[i];
            put(name, 0, v.size(), v.data(), v.size()*4);
        }
        void putI32(const std::string& name, const int32_t x)
        {
            put(name, 0, 1, &x, 4);
        }
        void putF64(const std::string& name, const scalarField& s)
        {
            put(name, 1, s.size(), s.begin(), size_t(s.size())*8);
        }
        void putStr(const std::string& name, const std::string& s)
        {
            put(name, 2, s.size(), s.data(), s.size());
        }
        bool write(const std::string& path)
        {
            FILE* f = fopen(path.c_str(), "wb");
            if (!f) return false;
            fwrite("B2LS0001", 1, 8, f);
            int64_t n = entries_.size();
            fwrite(&n, 8, 1, f);
            for (size_t i = 0; i < entries_.size(); i++)
            {
                int64_t nameLen = entries_[i].first.size();
                fwrite(&nameLen, 8, 1, f);
                fwrite(entries_[i].first.data(), 1, nameLen, f);
                fwrite(&types_[i].first, 8, 1, f);
                fwrite(&types_[i].second, 8, 1, f);
                if (entries_[i].second.size())
                {
                    fwrite(entries_[i].second.data(), 1, entries_[i].second.size(), f);
                }
            }
            fclose(f);
            return true;
        }
    };

public:

    TypeName("B200dump");

    B200dump
    (
        const word& fieldName,
        const lduMatrix& matrix,
        const Field<Field<scalar>>& interfaceBouCoeffs,
        const Field<Field<scalar>>& interfaceIntCoeffs,
        const lduInterfaceFieldPtrsList& interfaces,
        const dictionary& solverControls
    )
    :
        lduMatrix::solver
        (
            fieldName,
            matrix,
            interfaceBouCoeffs,
            interfaceIntCoeffs,
            interfaces,
            solverControls
        )
    {}

    virtual solverPerformance solve
    (
        scalarField& psi,
        const scalarField& source,
        const direction cmpt = 0
    ) const
    {
        static label nDumps = 0;

        const lduAddressing& addr = matrix_.lduAddr();
        writer w;
        w.putI32("nCells", int32_t(addr.size()));
        w.putI32("lower", addr.lowerAddr());
        w.putI32("upper", addr.upperAddr());
        w.putF64("diag", matrix_.diag());
        w.putF64("upperCoeffs", matrix_.upper());
        if (matrix_.asymmetric())
        {
            w.putF64("lowerCoeffs", matrix_.lower());
        }
        w.putF64("source", source);
        w.putF64("psi0", psi);

        // coupled patches, numbered among themselves as b200ls_mesh_create expects
        std::vector<label> coupled;
        forAll(interfaces_, patchi)
        {
            if (interfaces_.set(patchi)) coupled.push_back(patchi);
        }
        w.putI32("nIfaces", int32_t(coupled.size()));
        for (size_t k = 0; k < coupled.size(); k++)
        {
            const label patchi = coupled[k];
            const lduInterface& li = interfaces_[patchi].interface();
            const std::string key = "iface." + std::to_string(k);
            w.putI32(key + ".faceCells", li.faceCells());
            w.putF64(key + ".bouCoeffs", interfaceBouCoeffs_[patchi]);
            w.putF64(key + ".intCoeffs", interfaceIntCoeffs_[patchi]);
            if (isA<cyclicLduInterface>(li))
            {
                const label nbr =
                    refCast<const cyclicLduInterface>(li).nbrPatchIndex();
                int32_t nbrK = -1;
                for (size_t j = 0; j < coupled.size(); j++)
                {
                    if (coupled[j] == nbr) nbrK = j;
                }
                w.putI32(key + ".nbrPatch", nbrK);
            }
            else if (isA<processorLduInterface>(li))
            {
                w.putI32
                (
                    key + ".neighbProcNo",
                    int32_t
                    (
                        refCast<const processorLduInterface>(li).neighbProcNo()
                    )
                );
            }
        }

        // the solve as the delegate will see it
        dictionary d(controlDict_);
        const word delegate(controlDict_.lookup("delegate"));
        d.set("solver", delegate);
        d.remove("delegate");
        d.remove("dumpFile");
        {
            OStringStream os;
            d.write(os, false);
            w.putStr("solve.0.dict", os.str());
        }

        const fileName base
        (
            controlDict_.lookupOrDefault<fileName>("dumpFile", fieldName_)
        );
        std::string path =
            std::string(base) + "." + std::to_string(nDumps++);
        if (Pstream::parRun())
        {
            path += ".proc" + std::to_string(Pstream::myProcNo());
        }
        path += ".b2ls";
        if (!w.write(path))
        {
            FatalErrorInFunction
                << "cannot write " << path << exit(FatalError);
        }

        return lduMatrix::solver::New
        (
            fieldName_,
            matrix_,
            interfaceBouCoeffs_,
            interfaceIntCoeffs_,
            interfaces_,
            d
        )->solve(psi, source, cmpt);
    }
};


// ---------------------------------------------------------------------------
// run-time selection
// ---------------------------------------------------------------------------

defineTypeNameAndDebug(B200PCG, 0);
lduMatrix::solver::addsymMatrixConstructorToTable<B200PCG>
    addB200PCGSymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200PBiCGStab, 0);
lduMatrix::solver::addsymMatrixConstructorToTable<B200PBiCGStab>
    addB200PBiCGStabSymMatrixConstructorToTable_;
lduMatrix::solver::addasymMatrixConstructorToTable<B200PBiCGStab>
    addB200PBiCGStabAsymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200smoothSolver, 0);
lduMatrix::solver::addsymMatrixConstructorToTable<B200smoothSolver>
    addB200smoothSolverSymMatrixConstructorToTable_;
lduMatrix::solver::addasymMatrixConstructorToTable<B200smoothSolver>
    addB200smoothSolverAsymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200GAMG, 0);
lduMatrix::solver::addsymMatrixConstructorToTable<B200GAMG>
    addB200GAMGSymMatrixConstructorToTable_;
lduMatrix::solver::addasymMatrixConstructorToTable<B200GAMG>
    addB200GAMGAsymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200dump, 0);
lduMatrix::solver::addsymMatrixConstructorToTable<B200dump>
    addB200dumpSymMatrixConstructorToTable_;
lduMatrix::solver::addasymMatrixConstructorToTable<B200dump>
    addB200dumpAsymMatrixConstructorToTable_;

// preconditioners and smoothers register in the tables the reference's own
// DIC/DILU/GaussSeidel use (DICPreconditioner.C:34-36, DILUPreconditioner.C:34-36,
// GaussSeidelSmoother.C:34-38, DICSmoother.C:34-35, DILUSmoother.C:34-35)
defineTypeNameAndDebug(B200DICPreconditioner, 0);
lduMatrix::preconditioner::addsymMatrixConstructorToTable<B200DICPreconditioner>
    addB200DICPreconditionerSymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200DILUPreconditioner, 0);
lduMatrix::preconditioner::addasymMatrixConstructorToTable<B200DILUPreconditioner>
    addB200DILUPreconditionerAsymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200GaussSeidelSmoother, 0);
lduMatrix::smoother::addsymMatrixConstructorToTable<B200GaussSeidelSmoother>
    addB200GaussSeidelSmootherSymMatrixConstructorToTable_;
lduMatrix::smoother::addasymMatrixConstructorToTable<B200GaussSeidelSmoother>
    addB200GaussSeidelSmootherAsymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200symGaussSeidelSmoother, 0);
lduMatrix::smoother::addsymMatrixConstructorToTable<B200symGaussSeidelSmoother>
    addB200symGaussSeidelSmootherSymMatrixConstructorToTable_;
lduMatrix::smoother::addasymMatrixConstructorToTable<B200symGaussSeidelSmoother>
    addB200symGaussSeidelSmootherAsymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200DICSmoother, 0);
lduMatrix::smoother::addsymMatrixConstructorToTable<B200DICSmoother>
    addB200DICSmootherSymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200DILUSmoother, 0);
lduMatrix::smoother::addasymMatrixConstructorToTable<B200DILUSmoother>
    addB200DILUSmootherAsymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200DICGaussSeidelSmoother, 0);
lduMatrix::smoother::addsymMatrixConstructorToTable<B200DICGaussSeidelSmoother>
    addB200DICGaussSeidelSmootherSymMatrixConstructorToTable_;

defineTypeNameAndDebug(B200DILUGaussSeidelSmoother, 0);
lduMatrix::smoother::addasymMatrixConstructorToTable<B200DILUGaussSeidelSmoother>
    addB200DILUGaussSeidelSmootherAsymMatrixConstructorToTable_;

} // End namespace Foam
