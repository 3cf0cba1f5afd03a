// Matrix object (coefficients resident in HBM in the native layout) and the solver drivers.
#pragma once

#include "../../include/b200ls.h"
#include "device.cuh"

namespace b200ls {

// Coefficients of one level in the native layout
void arenaRelease(std::vector<std::pair<char*, size_t>>& blocks);

struct MatLevel {
    MatLevel() = default;
    MatLevel(MatLevel&&) = default;
    MatLevel& operator=(MatLevel&&) = default;
    ~MatLevel() { arenaRelease(arenaBlocks); }
    std::vector<std::pair<char*, size_t>> arenaBlocks;   // P2P arena blocks of this level (returned on destruction)
    DevBuf<double> diag;
    DevBuf<double> vals;            // [Uval (nFaces) | Lval (nFaces)]
    DevBuf<double> rD;              // DIC/DILU reciprocal diagonal
    bool rDValid = false;
    // pencil levels: coefficient planes [slot][position] (pencil.cuh k_pencil_planes / k_pencil_pack)
    DevBuf<double> pcL, pcLu, pcU;  // lower-side (K-, J-, I-) lower / upper coefficients, upper-side (I+, J+, K+)
    DevBuf<double> ptL, ptU;        // rD*lower in forward order, rD*upper in backward order (valid with rD)
    bool pPlanesValid = false;
    int rDKind = -1;                // which preconditioner rD belongs to (DIC/DILU vs diagonal)
    DevBuf<double> dWork;           // factorisation scratch (pre-reciprocal diagonal, sentinel protocol)
    // interfaces: coefficients, send/recv buffers
    std::vector<DevBuf<double>> bou, inn, sendBuf, recvBuf;
    DevBuf<unsigned char> ifaceViews;   // IfaceView[nIfaces] on device
    // P2P halos: receive buffers/flags live in this rank's IPC arena, the neighbour's are mapped pointers
    bool p2pReady = false;
    bool p2pTried = false;          // the (collective) P2P set-up of this level has run
    std::vector<double*> p2pRemoteRecv;                 // neighbour's receive buffer (2 parities)
    std::vector<unsigned long long*> p2pRemoteFlag;     // neighbour's epoch flag
    std::vector<double*> p2pLocalRecv;
    std::vector<unsigned long long*> p2pLocalFlag;
    DevBuf<unsigned int> p2pTickets;                    // one last-block ticket per interface
    unsigned long long haloEpoch = 0;
    // fused Gauss-Seidel sweeps across processor patches (setupCoupledGS)
    bool gsCoupled = false;
    int gsLag = 0;
    unsigned long long gsEpoch = 0;
    std::vector<double*> gsSlotLocal, gsSlotRemote;
    DevBuf<unsigned char> gsViews[2];                   // CoupledView[nIfaces] per call parity
    // level work vectors (GAMG): correction, source, scratch
    DevBuf<double> corr, src, tmpA, tmpB, tmpC;
    DevBuf<double> gsBufs;          // intermediate iterates of the fused multi-sweep Gauss-Seidel kernel
    bool tmpASentinel = false;      // tmpA is known to be all-sentinel
    double* Uval() { return vals.p; }
    double* Lval(int nFaces) { return vals.p + nFaces; }
};

struct Vec {
    DevBuf<double> buf;
    double* p() { return buf.p; }
};

}  // namespace b200ls

struct b200ls_matrix_s {
    b200ls_mesh_s* mesh = nullptr;
    int meshGeneration = -1;
    bool symmetric = true;
    bool valuesSet = false;
    bool coarseValid = false;       // coarse-level matrices match the current coefficients
    bool hasFingerprint = false;    // b200ls_matrix_set_if_changed: fingerprint of the coefficients held
    unsigned long long fingerprint = 0;
    std::vector<b200ls::MatLevel> levels;
    std::map<std::string, b200ls::Vec> vecs;     // named finest-level work vectors (position order)
    b200ls::DevBuf<double> stageA, stageB;       // cell-order staging (H2D / D2H)
    b200ls::DevBuf<double> scalars;              // device scalars of the Krylov loops
    b200ls::DevBuf<double> coarsestWork;         // scratch of the single-thread coarsest solve
    // gathered coarsest level (CoarsestGather): this rank's padded coefficient / source block and all ranks' blocks
    b200ls::DevBuf<double> gSendCoef, gCoef, gSendSrc, gSrc;
    double* vec(const std::string& name);
};

namespace b200ls {

void ensureDeviceMesh(b200ls_mesh_s* mesh);
// copies between (pageable) host memory and the device, ordered on the solver stream (capi.cu: CopyPool)
void h2dBytes(void* dev, const void* host, size_t bytes);
void d2hBytes(void* host, const void* dev, size_t bytes);
void matrixSet(b200ls_matrix_s* m, const double* diag, const double* upper, const double* lower,
               const double* const* bou, const double* const* inn, bool devicePointers = false);
void opAmul(b200ls_matrix_s* m, int level, double* out, const double* x);
void opResidual(b200ls_matrix_s* m, int level, double* out, const double* x, const double* b);
void opSumA(b200ls_matrix_s* m, int level, double* out);
void ensureFactor(b200ls_matrix_s* m, int level, int precond);
// dotOut != nullptr (DIC/DILU only): also compute wA.rA into dotOut[0] inside the backward sweep
void opPrecondition(b200ls_matrix_s* m, int level, int precond, double* wA, const double* rA,
                    double* dotOut = nullptr);
void opSmooth(b200ls_matrix_s* m, int level, int smoother, double*& psi, double*& spare, const double* source,
              int nSweeps);
void solveDev(b200ls_matrix_s* m, const b200ls_controls& c, double* psiCell, const double* sourceCell,
              b200ls_perf* perf);
void toPositions(b200ls_matrix_s* m, int level, double* outPos, const double* inCell);
void toCells(b200ls_matrix_s* m, int level, double* outCell, const double* inPos);
void checkSweepError();

}  // namespace b200ls
