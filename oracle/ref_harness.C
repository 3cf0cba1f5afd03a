/*---------------------------------------------------------------------------*\
  oracle/ref_harness.C  --  TEST INFRASTRUCTURE ONLY (never on the product path)

  Drives the UNMODIFIED reference solver stack (oracle/_ref/libOpenFOAM.so, built
  by oracle/build_ref.py from /root/reference/src) on an LDU system read from a
  "B2LS" case file and writes every result the parity tests compare against:

    * lduAddressing::losortAddr / ownerStartAddr / losortStartAddr
      (src/OpenFOAM/matrices/lduMatrix/lduAddressing/lduAddressing.C:32-170)
    * lduMatrix::Amul / residual / sumA   (lduMatrix/lduMatrixATmul.C:34-280)
    * DIC / DILU calcReciprocalD + precondition
    * GaussSeidel / DIC / DILU smoother sweeps (via lduMatrix::smoother::New)
    * lduMatrix::solver::New(...)->solve(...) for any fvSolution-style dictionary,
      repeated with maxIter = 1..k to obtain the exact residual history
    * GAMGAgglomeration levels (restrictAddressing, faceRestrictAddressing,
      faceFlipMap, coarse lower/upper) for the faceAreaPair agglomerator, which is
      restated here on top of pairGAMGAgglomeration exactly as
      src/finiteVolume/.../faceAreaPairGAMGAgglomeration.C:55-108 does, but taking the
      face weights from the case file instead of an fvMesh.

    * optional cyclic (same-process) coupled patches: the fine-level lduInterface /
      lduInterfaceField pair is supplied here (harnessCyclicInterface[Field], the
      lduPrimitiveMesh analogue of finiteVolume's cyclicFvPatch / cyclicFvPatchField.C:150-174);
      every coarse level then uses the reference's own cyclicGAMGInterface[Field], and all
      callers (Amul/residual/sumA/smoothers/solvers/GAMG) are the reference's.  This is the
      serial pin for the interface code that processor patches share.

    * `ref_harness --polymesh <caseDir> <out.b2ls>`: reads constant/polyMesh of a real case with the reference's
      polyMesh and writes what the synthetic generators need to build p-/U-like systems on it: LDU addressing
      (faceOwner/faceNeighbour of the internal faces), face area vectors, cell and face centres (the reference's own
      geometry engine, meshes/primitiveMesh/primitiveMeshFaceCentresAndAreas.C, ...CellCentresAndVols.C), cell
      volumes and the boundary patches (name, type, start, size, faceCells).

  One mesh per process: pairGAMGAgglomeration::forward_ is a process-global static.

  File formats are documented in openfoam-dev_b200/ldu_io.py.
\*---------------------------------------------------------------------------*/

#include "Time.H"
#include "lduPrimitiveMesh.H"
#include "lduMatrix.H"
#include "pairGAMGAgglomeration.H"
#include "GAMGAgglomeration.H"
#include "GAMGInterface.H"
#include "cyclicLduInterface.H"
#include "cyclicLduInterfaceField.H"
#include "lduInterfaceField.H"
#include "DICPreconditioner.H"
#include "DILUPreconditioner.H"
#include "addToRunTimeSelectionTable.H"
#include "IStringStream.H"
#include "OSspecific.H"
#include "clockTime.H"
#include "dlLibraryTable.H"
#include "polyMesh.H"

#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <map>

using namespace Foam;

// ---------------------------------------------------------------------------
// case-file globals (the harness handles exactly one mesh per process)
// ---------------------------------------------------------------------------

static std::vector<double> g_faceWeights;

namespace Foam
{

//- lduPrimitiveMesh with an objectRegistry so that GAMGAgglomeration::New can
//  cache itself (lduMesh::db() is NotImplemented in the base class,
//  meshes/lduMesh/lduMesh.C:40-45)
class harnessLduMesh
:
    public lduPrimitiveMesh
{
    const Time& runTime_;

public:

    harnessLduMesh
    (
        const Time& runTime,
        const label nCells,
        labelList& l,
        labelList& u
    )
    :
        lduPrimitiveMesh(nCells, l, u, 0, true),
        runTime_(runTime)
    {}

    virtual const objectRegistry& db() const
    {
        return runTime_;
    }
};



//- Fine-level cyclic coupled patch on an lduPrimitiveMesh (the role cyclicFvPatch plays on an fvMesh):
//  face i of this patch is coupled to face i of patch nbrIndex_ of the same mesh.
class harnessCyclicInterface
:
    public lduInterface,
    public cyclicLduInterface
{
    const label index_;
    const label nbrIndex_;
    const labelList faceCells_;
    const UPtrList<harnessCyclicInterface>& all_;
    const transformer transform_;

public:

    TypeName("cyclic");

    harnessCyclicInterface
    (
        const label index,
        const label nbrIndex,
        const labelList& faceCells,
        const UPtrList<harnessCyclicInterface>& all
    )
    :
        index_(index),
        nbrIndex_(nbrIndex),
        faceCells_(faceCells),
        all_(all),
        transform_()
    {}

    virtual const labelUList& faceCells() const
    {
        return faceCells_;
    }

    virtual tmp<labelField> interfaceInternalField(const labelUList& iF) const
    {
        tmp<labelField> t(new labelField(faceCells_.size()));
        labelField& f = t.ref();
        forAll(f, i)
        {
            f[i] = iF[faceCells_[i]];
        }
        return t;
    }

    virtual tmp<labelField> internalFieldTransfer
    (
        const Pstream::commsTypes,
        const labelUList& iF
    ) const
    {
        return all_[nbrIndex_].interfaceInternalField(iF);
    }

    virtual label nbrPatchIndex() const
    {
        return nbrIndex_;
    }

    virtual bool owner() const
    {
        return index_ < nbrIndex_;
    }

    virtual const cyclicLduInterface& nbrPatch() const
    {
        return all_[nbrIndex_];
    }

    virtual const transformer& transform() const
    {
        return transform_;
    }
};

defineTypeName(harnessCyclicInterface);


//- Fine-level cyclic interface field, scalar (rank 0, no transformation)
class harnessCyclicInterfaceField
:
    public lduInterfaceField,
    public cyclicLduInterfaceField
{
    const harnessCyclicInterface& patch_;

public:

    TypeName("cyclic");

    harnessCyclicInterfaceField(const harnessCyclicInterface& p)
    :
        lduInterfaceField(p),
        patch_(p)
    {}

    virtual const transformer& transform() const
    {
        return patch_.transform();
    }

    virtual int rank() const
    {
        return 0;
    }

    virtual void updateInterfaceMatrix
    (
        scalarField& result,
        const scalarField& psiInternal,
        const scalarField& coeffs,
        const direction,
        const Pstream::commsTypes
    ) const
    {
        const labelUList& nbrFaceCells =
            dynamic_cast<const harnessCyclicInterface&>(patch_.nbrPatch())
           .faceCells();
        const labelUList& faceCells = patch_.faceCells();
        scalarField pnf(psiInternal, nbrFaceCells);
        forAll(faceCells, i)
        {
            result[faceCells[i]] -= coeffs[i]*pnf[i];
        }
    }
};

defineTypeName(harnessCyclicInterfaceField);

//- faceAreaPair restated on pairGAMGAgglomeration with file-supplied weights
class harnessFaceAreaPairAgglomeration
:
    public pairGAMGAgglomeration
{
public:

    TypeName("faceAreaPair");

    harnessFaceAreaPairAgglomeration
    (
        const lduMesh& mesh,
        const dictionary& controlDict
    )
    :
        pairGAMGAgglomeration(mesh, controlDict)
    {
        scalarField w(g_faceWeights.size());
        forAll(w, i)
        {
            w[i] = g_faceWeights[i];
        }
        agglomerate(mesh, w);
    }
};

defineTypeNameAndDebug(harnessFaceAreaPairAgglomeration, 0);
addToRunTimeSelectionTable
(
    GAMGAgglomeration,
    harnessFaceAreaPairAgglomeration,
    lduMesh
);

}


// ---------------------------------------------------------------------------
// B2LS container IO: a flat list of named arrays
//   magic "B2LS0001"; int64 nEntries; per entry:
//   int64 nameLen; name bytes; int64 dtype (0=i32, 1=f64, 2=u8); int64 count; data
// ---------------------------------------------------------------------------

struct Entry
{
    int64_t dtype;
    std::vector<char> bytes;
    int64_t count;
};

typedef std::map<std::string, Entry> Container;

static size_t dtypeSize(int64_t dt)
{
    return dt == 0 ? 4 : dt == 1 ? 8 : 1;
}

static Container readContainer(const char* path)
{
    Container c;
    FILE* f = fopen(path, "rb");
    if (!f)
    {
        fprintf(stderr, "cannot open %s\n", path);
        exit(2);
    }
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "B2LS0001", 8) != 0)
    {
        fprintf(stderr, "bad magic in %s\n", path);
        exit(2);
    }
    int64_t n;
    if (fread(&n, 8, 1, f) != 1) exit(2);
    for (int64_t i = 0; i < n; i++)
    {
        int64_t nameLen;
        if (fread(&nameLen, 8, 1, f) != 1) exit(2);
        std::string name(nameLen, ' ');
        if (fread(&name[0], 1, nameLen, f) != size_t(nameLen)) exit(2);
        Entry e;
        if (fread(&e.dtype, 8, 1, f) != 1) exit(2);
        if (fread(&e.count, 8, 1, f) != 1) exit(2);
        e.bytes.resize(e.count*dtypeSize(e.dtype));
        if (e.bytes.size() && fread(e.bytes.data(), 1, e.bytes.size(), f) != e.bytes.size())
        {
            exit(2);
        }
        c[name] = e;
    }
    fclose(f);
    return c;
}

static void writeContainer(const char* path, const Container& c)
{
    FILE* f = fopen(path, "wb");
    if (!f)
    {
        fprintf(stderr, "cannot write %s\n", path);
        exit(2);
    }
    fwrite("B2LS0001", 1, 8, f);
    int64_t n = c.size();
    fwrite(&n, 8, 1, f);
    for (Container::const_iterator it = c.begin(); it != c.end(); ++it)
    {
        int64_t nameLen = it->first.size();
        fwrite(&nameLen, 8, 1, f);
        fwrite(it->first.data(), 1, nameLen, f);
        fwrite(&it->second.dtype, 8, 1, f);
        fwrite(&it->second.count, 8, 1, f);
        if (it->second.bytes.size())
        {
            fwrite(it->second.bytes.data(), 1, it->second.bytes.size(), f);
        }
    }
    fclose(f);
}

static bool has(const Container& c, const std::string& k)
{
    return c.find(k) != c.end();
}

static const int32_t* i32(const Container& c, const std::string& k)
{
    return reinterpret_cast<const int32_t*>(c.at(k).bytes.data());
}

static const double* f64(const Container& c, const std::string& k)
{
    return reinterpret_cast<const double*>(c.at(k).bytes.data());
}

static std::string str(const Container& c, const std::string& k)
{
    const Entry& e = c.at(k);
    return std::string(e.bytes.begin(), e.bytes.end());
}

static void putI32(Container& c, const std::string& k, const labelUList& l)
{
    Entry e;
    e.dtype = 0;
    e.count = l.size();
    e.bytes.resize(4*l.size());
    if (l.size()) memcpy(e.bytes.data(), l.begin(), 4*l.size());
    c[k] = e;
}

static void putF64(Container& c, const std::string& k, const UList<scalar>& l)
{
    Entry e;
    e.dtype = 1;
    e.count = l.size();
    e.bytes.resize(8*l.size());
    if (l.size()) memcpy(e.bytes.data(), l.begin(), 8*l.size());
    c[k] = e;
}

static void putF64(Container& c, const std::string& k, const std::vector<double>& l)
{
    Entry e;
    e.dtype = 1;
    e.count = l.size();
    e.bytes.resize(8*l.size());
    if (l.size()) memcpy(e.bytes.data(), l.data(), 8*l.size());
    c[k] = e;
}

static void putU8(Container& c, const std::string& k, const boolList& l)
{
    Entry e;
    e.dtype = 2;
    e.count = l.size();
    e.bytes.resize(l.size());
    forAll(l, i)
    {
        e.bytes[i] = l[i] ? 1 : 0;
    }
    c[k] = e;
}

static scalarField toField(const Container& c, const std::string& k)
{
    const Entry& e = c.at(k);
    scalarField f(e.count);
    if (e.count) memcpy(f.begin(), e.bytes.data(), 8*e.count);
    return f;
}


// ---------------------------------------------------------------------------


static int dumpPolyMesh(const char* caseDir, const char* outPath)
{
    fileName casePath(caseDir);
    Time runTime
    (
        Time::controlDictName,
        casePath.path(),
        casePath.name(),
        false
    );
    polyMesh mesh
    (
        IOobject
        (
            polyMesh::defaultRegion,
            runTime.name(),
            runTime,
            IOobject::MUST_READ
        )
    );

    Container out;
    const label nInt = mesh.nInternalFaces();
    labelList sizes(3);
    sizes[0] = mesh.nCells();
    sizes[1] = nInt;
    sizes[2] = mesh.nFaces();
    putI32(out, "sizes", sizes);
    putI32(out, "lower", labelList(SubList<label>(mesh.faceOwner(), nInt)));
    putI32(out, "upper", labelList(mesh.faceNeighbour()));
    putI32(out, "faceOwner", mesh.faceOwner());

    const vectorField& Sf = mesh.faceAreas();
    const vectorField& Cf = mesh.faceCentres();
    const vectorField& C = mesh.cellCentres();
    scalarField buf(3*Sf.size());
    forAll(Sf, i)
    {
        buf[3*i] = Sf[i].x(); buf[3*i + 1] = Sf[i].y(); buf[3*i + 2] = Sf[i].z();
    }
    putF64(out, "faceAreas", buf);
    forAll(Cf, i)
    {
        buf[3*i] = Cf[i].x(); buf[3*i + 1] = Cf[i].y(); buf[3*i + 2] = Cf[i].z();
    }
    putF64(out, "faceCentres", buf);
    scalarField cbuf(3*C.size());
    forAll(C, i)
    {
        cbuf[3*i] = C[i].x(); cbuf[3*i + 1] = C[i].y(); cbuf[3*i + 2] = C[i].z();
    }
    putF64(out, "cellCentres", cbuf);
    putF64(out, "cellVolumes", mesh.cellVolumes());

    const polyBoundaryMesh& bm = mesh.boundary();
    labelList nPatches(1, bm.size());
    putI32(out, "nPatches", nPatches);
    forAll(bm, pi)
    {
        const std::string k = "patch." + std::to_string(pi);
        labelList ss(2);
        ss[0] = bm[pi].start();
        ss[1] = bm[pi].size();
        putI32(out, k + ".startSize", ss);
        Entry e;
        const std::string desc = std::string(bm[pi].name()) + " " + std::string(bm[pi].type());
        e.dtype = 2;
        e.count = desc.size();
        e.bytes.assign(desc.begin(), desc.end());
        out[k + ".nameType"] = e;
    }
    writeContainer(outPath, out);
    return 0;
}

int main(int argc, char* argv[])
{
    if (argc == 4 && std::string(argv[1]) == "--polymesh")
    {
        return dumpPolyMesh(argv[2], argv[3]);
    }

    if (argc < 3)
    {
        fprintf
        (
            stderr,
            "usage: ref_harness <case.b2ls> <out.b2ls> [scratchCaseDir]\n"
        );
        return 2;
    }

    Container in = readContainer(argv[1]);
    Container out;

    // Scratch case directory with a minimal system/controlDict for the Time
    // object that backs lduMesh::db()
    std::string scratch = argc > 3 ? argv[3] : "/tmp/b200ls_ref_case";
    mkDir(fileName(scratch)/"system");
    mkDir(fileName(scratch)/"constant");
    {
        FILE* f = fopen((scratch + "/system/controlDict").c_str(), "w");
        fprintf
        (
            f,
            "FoamFile{format ascii; class dictionary; object controlDict;}\n"
            "application harness;\nstartFrom startTime;\nstartTime 0;\n"
            "stopAt endTime;\nendTime 1;\ndeltaT 1;\nwriteControl timeStep;\n"
            "writeInterval 1000000;\n"
        );
        fclose(f);
    }

    fileName scratchPath(scratch);
    Time runTime
    (
        Time::controlDictName,
        scratchPath.path(),
        scratchPath.name(),
        false
    );

    // Optional plugin libraries (the drop-in boundary: controlDict `libs`)
    if (has(in, "libs"))
    {
        IStringStream is("libs (" + str(in, "libs") + ");");
        dictionary libsDict(is);
        libs.open(libsDict, "libs");
    }

    const label nCells = i32(in, "nCells")[0];
    const label nFaces = in.at("lower").count;

    labelList l(nFaces), u(nFaces);
    for (label i = 0; i < nFaces; i++)
    {
        l[i] = i32(in, "lower")[i];
        u[i] = i32(in, "upper")[i];
    }

    if (has(in, "faceWeights"))
    {
        g_faceWeights.assign
        (
            f64(in, "faceWeights"),
            f64(in, "faceWeights") + in.at("faceWeights").count
        );
    }

    harnessLduMesh mesh(runTime, nCells, l, u);
    // cyclic coupled patches: "nIfaces", "iface.<i>.{faceCells,nbrPatch,bouCoeffs,intCoeffs}"
    const label nIfaces = has(in, "nIfaces") ? i32(in, "nIfaces")[0] : 0;
    // owned by the mesh once added (lduPrimitiveMesh::addInterfaces, lduPrimitiveMesh.C:219-237)
    UPtrList<harnessCyclicInterface> cyclicPatches(nIfaces);
    PtrList<harnessCyclicInterfaceField> cyclicFields(nIfaces);
    lduInterfacePtrsList meshInterfaces(nIfaces);
    lduInterfaceFieldPtrsList ifaceFields(nIfaces);
    Field<Field<scalar>> bouCoeffs(nIfaces), intCoeffs(nIfaces);
    for (label i = 0; i < nIfaces; i++)
    {
        const std::string k = "iface." + std::to_string(i);
        const label n = in.at(k + ".faceCells").count;
        labelList fc(n);
        for (label j = 0; j < n; j++)
        {
            fc[j] = i32(in, k + ".faceCells")[j];
        }
        cyclicPatches.set
        (
            i,
            new harnessCyclicInterface
            (
                i, i32(in, k + ".nbrPatch")[0], fc, cyclicPatches
            )
        );
        meshInterfaces.set(i, &cyclicPatches[i]);
        cyclicFields.set(i, new harnessCyclicInterfaceField(cyclicPatches[i]));
        ifaceFields.set(i, &cyclicFields[i]);
        bouCoeffs[i] = toField(in, k + ".bouCoeffs");
        intCoeffs[i] = toField(in, k + ".intCoeffs");
    }
    mesh.addInterfaces(meshInterfaces, lduSchedule());

    // --- integer addressing -------------------------------------------------
    putI32(out, "losort", mesh.lduAddr().losortAddr());
    putI32(out, "ownerStart", mesh.lduAddr().ownerStartAddr());
    putI32(out, "losortStart", mesh.lduAddr().losortStartAddr());

    lduMatrix A(mesh);
    if (has(in, "diag"))
    {
        A.diag() = toField(in, "diag");
        A.upper() = toField(in, "upperCoeffs");
        if (has(in, "lowerCoeffs"))
        {
            A.lower() = toField(in, "lowerCoeffs");
        }
    }


    // --- operators ------------------------------------------------------------
    if (has(in, "x"))
    {
        const scalarField x(toField(in, "x"));
        scalarField Ax(nCells);
        A.Amul(Ax, x, bouCoeffs, ifaceFields, 0);
        putF64(out, "Amul", Ax);

        scalarField sA(nCells);
        A.sumA(sA, bouCoeffs, ifaceFields);
        putF64(out, "sumA", sA);

        if (has(in, "source"))
        {
            scalarField rA(nCells);
            A.residual(rA, x, toField(in, "source"), bouCoeffs, ifaceFields, 0);
            putF64(out, "residual", rA);
        }

        // DIC (symmetric) or DILU (asymmetric) factor + one application to x
        if (!A.asymmetric())
        {
            scalarField rD(A.diag());
            DICPreconditioner::calcReciprocalD(rD, A);
            putF64(out, "DIC_rD", rD);
        }
        else
        {
            scalarField rD(A.diag());
            DILUPreconditioner::calcReciprocalD(rD, A);
            putF64(out, "DILU_rD", rD);
        }
        {
            // precondition through the selection table, as PCG/PBiCGStab do
            IStringStream is
            (
                A.asymmetric()
              ? "solver PBiCGStab; preconditioner DILU;"
              : "solver PCG; preconditioner DIC;"
            );
            dictionary d(is);
            autoPtr<lduMatrix::solver> sol = lduMatrix::solver::New
            (
                "p", A, bouCoeffs, intCoeffs, ifaceFields, d
            );
            autoPtr<lduMatrix::preconditioner> pre =
                lduMatrix::preconditioner::New(sol(), d);
            scalarField wA(nCells);
            pre->precondition(wA, x, 0);
            putF64(out, "precondition", wA);
        }
    }

    // --- smoothers: entries "smooth.<i>.dict" (e.g. "smoother GaussSeidel;"),
    //     "smooth.<i>.nSweeps"; psi0 = x, rhs = source
    for (int i = 0; ; i++)
    {
        const std::string key = "smooth." + std::to_string(i);
        if (!has(in, key + ".dict")) break;

        IStringStream is(str(in, key + ".dict"));
        dictionary d(is);
        autoPtr<lduMatrix::smoother> sm = lduMatrix::smoother::New
        (
            "p", A, bouCoeffs, intCoeffs, ifaceFields, d
        );
        scalarField psi(toField(in, "x"));
        const scalarField b(toField(in, "source"));
        sm->smooth(psi, b, 0, i32(in, key + ".nSweeps")[0]);
        putF64(out, key + ".psi", psi);
    }

    // --- solves: entries "solve.<i>.dict", optional "solve.<i>.history" = k:
    //     re-run with maxIter=1..k (same tolerances) and store the final residual of each run
    for (int i = 0; ; i++)
    {
        const std::string key = "solve." + std::to_string(i);
        if (!has(in, key + ".dict")) break;

        const std::string dictStr = str(in, key + ".dict");
        const scalarField b(toField(in, "source"));
        const scalarField psi0
        (
            has(in, "psi0") ? toField(in, "psi0") : scalarField(nCells, 0.0)
        );

        IStringStream is(dictStr);
        dictionary d(is);

        scalarField psi(psi0);
        clockTime timer;
        solverPerformance perf = lduMatrix::solver::New
        (
            "p", A, bouCoeffs, intCoeffs, ifaceFields, d
        )->solve(psi, b);
        const double secs = timer.elapsedTime();

        std::vector<double> p(6);
        p[0] = perf.initialResidual();
        p[1] = perf.finalResidual();
        p[2] = perf.nIterations();
        p[3] = perf.converged();
        p[4] = perf.singular();
        p[5] = secs;
        putF64(out, key + ".perf", p);
        putF64(out, key + ".psi", psi);
        {
            Entry e;
            const std::string nm = perf.solverName();
            e.dtype = 2;
            e.count = nm.size();
            e.bytes.assign(nm.begin(), nm.end());
            out[key + ".solverName"] = e;
        }

        if (has(in, key + ".history"))
        {
            const label k = i32(in, key + ".history")[0];
            std::vector<double> hist;
            for (label it = 1; it <= k; it++)
            {
                dictionary dk(d);
                // tolerance/relTol are kept: a GAMG coarsest solver inherits them and must
                // still terminate.  Entries past convergence repeat the final residual.
                dk.set("maxIter", it);
                scalarField psik(psi0);
                solverPerformance pk = lduMatrix::solver::New
                (
                    "p", A, bouCoeffs, intCoeffs, ifaceFields, dk
                )->solve(psik, b);
                hist.push_back(pk.finalResidual());
            }
            putF64(out, key + ".historyResiduals", hist);
        }
    }

    // --- agglomeration dump: entry "agglomerate.dict" (GAMG controls) --------
    if (has(in, "agglomerate.dict"))
    {
        IStringStream is(str(in, "agglomerate.dict"));
        dictionary d(is);
        const GAMGAgglomeration& agg = GAMGAgglomeration::New(mesh, d);

        labelList nLevels(1, agg.size());
        putI32(out, "agg.nLevels", nLevels);

        for (label lev = 0; lev < agg.size(); lev++)
        {
            const std::string k = "agg." + std::to_string(lev);
            labelList sizes(2);
            sizes[0] = agg.nCells(lev);
            sizes[1] = agg.nFaces(lev);
            putI32(out, k + ".sizes", sizes);
            putI32(out, k + ".restrictAddressing", agg.restrictAddressing(lev));
            putI32
            (
                out,
                k + ".faceRestrictAddressing",
                agg.faceRestrictAddressing(lev)
            );
            putU8(out, k + ".faceFlipMap", agg.faceFlipMap(lev));
            const lduAddressing& ca = agg.meshLevel(lev + 1).lduAddr();
            putI32(out, k + ".coarseLower", ca.lowerAddr());
            putI32(out, k + ".coarseUpper", ca.upperAddr());

            const lduInterfacePtrsList ci(agg.meshLevel(lev + 1).interfaces());
            forAll(ci, pi)
            {
                if (!ci.set(pi)) continue;
                const GAMGInterface& gi = refCast<const GAMGInterface>(ci[pi]);
                const std::string ki = k + ".iface." + std::to_string(pi);
                putI32(out, ki + ".faceCells", gi.faceCells());
                putI32
                (
                    out,
                    ki + ".faceRestrictAddressing",
                    gi.faceRestrictAddressing()
                );
            }
        }
    }

    writeContainer(argv[2], out);

    return 0;
}
