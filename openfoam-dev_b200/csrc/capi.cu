// extern "C" surface of libb200ls.so (include/b200ls.h): argument checking, exception -> error-code
// translation, host<->device staging.  No compute lives here.
#include <dlfcn.h>

#include <cstring>
#include <string>

#include "kernels.cuh"
#include "solver.cuh"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

using namespace b200ls;

namespace {

thread_local std::string g_lastError;
bool g_forward = true;   // mirror of the reference's process-global pairGAMGAgglomeration::forward_

template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return 1;
    } catch (...) {
        g_lastError = "unknown error";
        return 1;
    }
}

template <class T>
T* sym(void* h, const char* name) {
    void* p = dlsym(h, name);
    if (!p) throw CudaError(std::string("NCCL symbol not found: ") + name);
    return reinterpret_cast<T*>(p);
}

void loadNccl(NcclApi& n) {
    if (n.handle) return;
    n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.handle) n.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.handle) throw CudaError(std::string("cannot load NCCL: ") + dlerror());
    n.GetUniqueId = sym<std::remove_pointer_t<decltype(n.GetUniqueId)>>(n.handle, "ncclGetUniqueId");
    n.CommInitRank = sym<std::remove_pointer_t<decltype(n.CommInitRank)>>(n.handle, "ncclCommInitRank");
    n.CommDestroy = sym<std::remove_pointer_t<decltype(n.CommDestroy)>>(n.handle, "ncclCommDestroy");
    n.AllReduce = sym<std::remove_pointer_t<decltype(n.AllReduce)>>(n.handle, "ncclAllReduce");
    n.AllGather = sym<std::remove_pointer_t<decltype(n.AllGather)>>(n.handle, "ncclAllGather");
    n.Send = sym<std::remove_pointer_t<decltype(n.Send)>>(n.handle, "ncclSend");
    n.Recv = sym<std::remove_pointer_t<decltype(n.Recv)>>(n.handle, "ncclRecv");
    n.GroupStart = sym<std::remove_pointer_t<decltype(n.GroupStart)>>(n.handle, "ncclGroupStart");
    n.GroupEnd = sym<std::remove_pointer_t<decltype(n.GroupEnd)>>(n.handle, "ncclGroupEnd");
    n.GetErrorString = sym<std::remove_pointer_t<decltype(n.GetErrorString)>>(n.handle, "ncclGetErrorString");
}

// Export this rank's P2P arena with CUDA IPC and map every peer's (same node, NVLink/NVSwitch).  Collective: either
// every rank ends up with P2P enabled or none does (then all traffic stays on NCCL).
void setupP2P(Context& c) {
    P2PState& P = c.p2p;
    if (P.enabled || c.nRanks < 2 || c.nRanks > kMaxRanks || getenv("B200LS_NO_P2P")) return;
    size_t bytes = size_t(256) << 20;
    if (const char* s = getenv("B200LS_P2P_ARENA_MB")) bytes = size_t(atol(s)) << 20;
    bool ok = true;
    if (cudaMalloc(&P.arena, bytes) != cudaSuccess) {
        cudaGetLastError();
        P.arena = nullptr;
        ok = false;
    }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) {
        B2_CUDA(cudaMemsetAsync(P.arena, 0, bytes, c.stream));
        if (cudaIpcGetMemHandle(&mine, P.arena) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
        }
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DevBuf<char> dMine, dAll;
    dMine.alloc(64);
    dAll.alloc(size_t(64) * c.nRanks);
    B2_CUDA(cudaMemcpyAsync(dMine.p, &mine, 64, cudaMemcpyHostToDevice, c.stream));
    if (c.nccl.AllGather(dMine.p, dAll.p, 64, ncclChar, c.comm, c.stream) != 0) throw CudaError("ncclAllGather failed");
    std::vector<cudaIpcMemHandle_t> all(c.nRanks);
    B2_CUDA(cudaMemcpyAsync(all.data(), dAll.p, size_t(64) * c.nRanks, cudaMemcpyDeviceToHost, c.stream));
    B2_CUDA(cudaStreamSynchronize(c.stream));
    P.view.rank = c.rank;
    P.view.nRanks = c.nRanks;
    for (int r = 0; r < c.nRanks && ok; r++) {
        if (r == c.rank) {
            P.view.peer[r] = P.arena;
            continue;
        }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            break;
        }
        P.opened.push_back(ptr);
        P.view.peer[r] = static_cast<char*>(ptr);
    }
    // agree: P2P only if it worked everywhere (this all-reduce is also the barrier after the arena memsets)
    DevBuf<double> flag;
    flag.alloc(1);
    const double bad = ok ? 0.0 : 1.0;
    B2_CUDA(cudaMemcpyAsync(flag.p, &bad, sizeof(double), cudaMemcpyHostToDevice, c.stream));
    if (c.nccl.AllReduce(flag.p, flag.p, 1, ncclDouble, ncclSum, c.comm, c.stream) != 0)
        throw CudaError("ncclAllReduce failed");
    double total = 0;
    B2_CUDA(cudaMemcpyAsync(&total, flag.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    B2_CUDA(cudaStreamSynchronize(c.stream));
    if (total != 0.0) {
        for (void* q : P.opened) cudaIpcCloseMemHandle(q);
        P.opened.clear();
        if (P.arena) cudaFree(P.arena);
        P.arena = nullptr;
        return;
    }
    P.arenaBytes = bytes;
    P.bump = kP2PReduceBytes;
    P.enabled = true;
}

}  // namespace

namespace b200ls {
// ---- staged copies between PAGEABLE host memory and the device ---------------------------------------------
// OpenFOAM's fields live in pageable memory.  A plain cudaMemcpy stages them through one pinned buffer with one host
// thread (~10 GB/s measured for the 100 MB of a 128^3 solve).  Here a few persistent worker threads copy 2 MB chunks (B200LS_COPY_CHUNK_KB)
// into their own pinned double buffers and issue the DMA on their own streams, so host memcpy and PCIe overlap and
// several cores feed the link.  B200LS_COPY_THREADS (default 4; 0 or 1 = plain cudaMemcpyAsync).
class CopyPool {
public:
    static size_t chunkBytes() {
        static const size_t v = getenv("B200LS_COPY_CHUNK_KB") ? size_t(atol(getenv("B200LS_COPY_CHUNK_KB"))) << 10 : size_t(2) << 20;
        return v;
    }
    static constexpr size_t kMinBytes = size_t(4) << 20;
    struct Lane {
        char* pinned[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaEvent_t done = nullptr;
        cudaStream_t stream = nullptr;
    };
    int nThreads = 0;
    std::vector<Lane> lanes;

    static CopyPool* get() {
        static CopyPool* pool = nullptr;   // never destroyed: the workers just end with the process
        static bool tried = false;
        if (!tried) {
            tried = true;
            int n = 4;
            if (const char* e = getenv("B200LS_COPY_THREADS")) n = atoi(e);
            const unsigned hw = std::thread::hardware_concurrency();
            if (hw && n > int(hw)) n = int(hw);
            if (n >= 2) pool = new CopyPool(n);
        }
        return pool;
    }

    // run fn(lane index) on every worker and wait
    void run(const std::function<void(int)>& fn) {
        std::unique_lock<std::mutex> lk(mu_);
        job_ = &fn;
        pending_ = nThreads;
        generation_++;
        cv_.notify_all();
        done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
        if (!error_.empty()) {
            std::string e = error_;
            error_.clear();
            throw CudaError(e);
        }
    }

private:
    explicit CopyPool(int n) : nThreads(n), lanes(n) {
        const int device = ctx().device;
        for (int t = 0; t < n; t++) {
            Lane& L = lanes[t];
            B2_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
            B2_CUDA(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
            for (int b = 0; b < 2; b++) {
                B2_CUDA(cudaMallocHost(&L.pinned[b], chunkBytes()));
                B2_CUDA(cudaEventCreateWithFlags(&L.ev[b], cudaEventDisableTiming));
            }
        }
        for (int t = 0; t < n; t++) std::thread([this, t, device] { worker(t, device); }).detach();
    }
    void worker(int t, int device) {
        cudaSetDevice(device);
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)>* fn = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                fn = job_;
            }
            std::string err;
            try {
                (*fn)(t);
            } catch (const std::exception& e) {
                err = e.what();
            }
            std::lock_guard<std::mutex> lk(mu_);
            if (!err.empty()) error_ = err;
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    unsigned long long generation_ = 0;
    int pending_ = 0;
    std::string error_;
};

// host (pageable or pinned) -> device, ordered like an async copy on the solver stream: later work on that stream sees
// the data.  The host buffer may be reused as soon as this returns.
void h2dBytes(void* dev, const void* host, size_t bytes) {
    if (!bytes) return;
    Context& c = ctx();
    CopyPool* pool = bytes >= CopyPool::kMinBytes ? CopyPool::get() : nullptr;
    if (!pool) {
        B2_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c.stream));
        return;
    }
    // the lanes' streams start after whatever the solver stream still does with `dev`
    cudaEvent_t start;
    B2_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    B2_CUDA(cudaEventRecord(start, c.stream));
    const size_t nChunks = (bytes + CopyPool::chunkBytes() - 1) / CopyPool::chunkBytes();
    const int T = pool->nThreads;
    pool->run([&](int t) {
        CopyPool::Lane& L = pool->lanes[t];
        B2_CUDA(cudaStreamWaitEvent(L.stream, start, 0));
        int b = 0;
        for (size_t k = t; k < nChunks; k += T, b ^= 1) {
            const size_t off = k * CopyPool::chunkBytes(), len = std::min(CopyPool::chunkBytes(), bytes - off);
            B2_CUDA(cudaEventSynchronize(L.ev[b]));   // the previous DMA out of this buffer has finished
            memcpy(L.pinned[b], static_cast<const char*>(host) + off, len);
            B2_CUDA(cudaMemcpyAsync(static_cast<char*>(dev) + off, L.pinned[b], len, cudaMemcpyHostToDevice, L.stream));
            B2_CUDA(cudaEventRecord(L.ev[b], L.stream));
        }
        B2_CUDA(cudaEventRecord(L.done, L.stream));
    });
    for (int t = 0; t < T; t++) B2_CUDA(cudaStreamWaitEvent(c.stream, pool->lanes[t].done, 0));
    cudaEventDestroy(start);
}

// device -> host (pageable or pinned); returns when the data is in `host`
void d2hBytes(void* host, const void* dev, size_t bytes) {
    if (!bytes) return;
    Context& c = ctx();
    CopyPool* pool = bytes >= CopyPool::kMinBytes ? CopyPool::get() : nullptr;
    if (!pool) {
        B2_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
        B2_CUDA(cudaStreamSynchronize(c.stream));
        return;
    }
    cudaEvent_t start;
    B2_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    B2_CUDA(cudaEventRecord(start, c.stream));
    const size_t nChunks = (bytes + CopyPool::chunkBytes() - 1) / CopyPool::chunkBytes();
    const int T = pool->nThreads;
    pool->run([&](int t) {
        CopyPool::Lane& L = pool->lanes[t];
        B2_CUDA(cudaStreamWaitEvent(L.stream, start, 0));
        int b = 0;
        long long prev = -1;   // chunk whose DMA into buffer b^1 is in flight
        for (size_t k = t; k < nChunks; k += T, b ^= 1) {
            const size_t off = k * CopyPool::chunkBytes(), len = std::min(CopyPool::chunkBytes(), bytes - off);
            B2_CUDA(cudaMemcpyAsync(L.pinned[b], static_cast<const char*>(dev) + off, len, cudaMemcpyDeviceToHost, L.stream));
            B2_CUDA(cudaEventRecord(L.ev[b], L.stream));
            if (prev >= 0) {
                const size_t poff = size_t(prev) * CopyPool::chunkBytes(), plen = std::min(CopyPool::chunkBytes(), bytes - poff);
                B2_CUDA(cudaEventSynchronize(L.ev[b ^ 1]));
                memcpy(static_cast<char*>(host) + poff, L.pinned[b ^ 1], plen);
            }
            prev = (long long)k;
        }
        if (prev >= 0) {
            const size_t poff = size_t(prev) * CopyPool::chunkBytes(), plen = std::min(CopyPool::chunkBytes(), bytes - poff);
            B2_CUDA(cudaEventSynchronize(L.ev[b ^ 1]));
            memcpy(static_cast<char*>(host) + poff, L.pinned[b ^ 1], plen);
        }
    });
    cudaEventDestroy(start);
}

}  // namespace b200ls

namespace {
// stage a host vector of level-0 size into `dev` (cell order)
void h2d(double* dev, const double* host, int n) { h2dBytes(dev, host, sizeof(double) * size_t(n)); }
void d2h(double* host, const double* dev, int n) {
    d2hBytes(host, dev, sizeof(double) * size_t(n));
    B2_CUDA(cudaStreamSynchronize(ctx().stream));
}

// ---- host-side communicator for multi-rank agglomeration ---------------------------------------------------
// Either caller-supplied callbacks (b200ls_set_host_comm: Pstream in the plugin, torch.distributed/gloo in the
// CPU tests) or NCCL on small staging buffers.
typedef void (*exchange_cb_t)(int32_t nIfaces, const int32_t* nbr, const int32_t* sizes, const int32_t* const* send,
                              int32_t* const* recv);
typedef int64_t (*sum_cb_t)(int64_t);
exchange_cb_t g_exchangeCb = nullptr;
sum_cb_t g_sumCb = nullptr;
int g_hostRank = 0, g_hostRanks = 1;
bool g_hostCommSet = false;

HostComm makeHostComm() {
    HostComm hc;
    if (g_hostCommSet) {
        hc.exchange = [](const std::vector<int32_t>& nbr, const std::vector<std::vector<int32_t>>& send,
                         std::vector<std::vector<int32_t>>& recv) {
            std::vector<int32_t> sizes(nbr.size());
            std::vector<const int32_t*> sp(nbr.size());
            std::vector<int32_t*> rp(nbr.size());
            for (size_t i = 0; i < nbr.size(); i++) {
                sizes[i] = int32_t(send[i].size());
                sp[i] = send[i].data();
                rp[i] = recv[i].data();
            }
            g_exchangeCb(int32_t(nbr.size()), nbr.data(), sizes.data(), sp.data(), rp.data());
        };
        hc.sum = [](int64_t v) { return g_sumCb(v); };
        return hc;
    }
    hc.exchange = [](const std::vector<int32_t>& nbr, const std::vector<std::vector<int32_t>>& send,
                     std::vector<std::vector<int32_t>>& recv) {
        Context& c = ctx();
        ensureInit();
        if (!c.comm) throw CudaError("multi-rank agglomeration needs b200ls_init with nRanks > 1");
        std::vector<DevBuf<int>> ds(nbr.size()), dr(nbr.size());
        for (size_t i = 0; i < nbr.size(); i++) {
            ds[i].upload(std::vector<int>(send[i].begin(), send[i].end()), c.stream);
            dr[i].alloc(send[i].size());
        }
        B2_CUDA(cudaStreamSynchronize(c.stream));
        c.nccl.GroupStart();
        for (size_t i = 0; i < nbr.size(); i++) {
            c.nccl.Send(ds[i].p, send[i].size(), ncclInt32, nbr[i], c.comm, c.stream);
            c.nccl.Recv(dr[i].p, send[i].size(), ncclInt32, nbr[i], c.comm, c.stream);
        }
        int r = c.nccl.GroupEnd();
        if (r != 0) throw CudaError(std::string("nccl restrictMap exchange: ") + c.nccl.GetErrorString((ncclResult_t)r));
        for (size_t i = 0; i < nbr.size(); i++) {
            if (!send[i].empty())
                B2_CUDA(cudaMemcpyAsync(recv[i].data(), dr[i].p, sizeof(int32_t) * send[i].size(),
                                        cudaMemcpyDeviceToHost, c.stream));
        }
        B2_CUDA(cudaStreamSynchronize(c.stream));
    };
    hc.sum = [](int64_t v) {
        Context& c = ctx();
        ensureInit();
        if (!c.comm) throw CudaError("multi-rank agglomeration needs b200ls_init with nRanks > 1");
        DevBuf<long long> d;
        d.alloc(1);
        long long h = v;
        B2_CUDA(cudaMemcpyAsync(d.p, &h, sizeof(h), cudaMemcpyHostToDevice, c.stream));
        int r = c.nccl.AllReduce(d.p, d.p, 1, ncclInt64, ncclSum, c.comm, c.stream);
        if (r != 0) throw CudaError("ncclAllReduce(int64) failed");
        B2_CUDA(cudaMemcpyAsync(&h, d.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
        B2_CUDA(cudaStreamSynchronize(c.stream));
        return int64_t(h);
    };
    return hc;
}

void requireValues(b200ls_matrix_t m) {
    if (!m || !m->valuesSet) throw CudaError("matrix coefficients not set (call b200ls_matrix_set)");
}

}  // namespace

extern "C" {

const char* b200ls_last_error(void) { return g_lastError.c_str(); }

int b200ls_device_available(void) {
    int n = 0;
    return (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0;
}

int b200ls_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200ls_nccl_unique_id(void* out128) {
    return guarded([&] {
        loadNccl(ctx().nccl);
        ncclUniqueId id;
        int r = ctx().nccl.GetUniqueId(&id);
        if (r != 0) throw CudaError("ncclGetUniqueId failed");
        static_assert(sizeof(id) == 128, "ncclUniqueId size");
        memcpy(out128, &id, 128);
    });
}

int b200ls_init(int device, const void* ncclUniqueId_, int rank, int nRanks) {
    return guarded([&] {
        Context& c = ctx();
        if (c.initialised && (c.device != device || c.nRanks != nRanks || c.rank != rank)) {
            throw CudaError("b200ls_init called twice with different arguments");
        }
        c.device = device;
        ensureInit();
        if (nRanks > 1 && !c.comm) {
            if (!ncclUniqueId_) throw CudaError("nRanks > 1 needs an ncclUniqueId");
            loadNccl(c.nccl);
            ncclUniqueId id;
            memcpy(&id, ncclUniqueId_, 128);
            int r = c.nccl.CommInitRank(&c.comm, nRanks, id, rank);
            if (r != 0) throw CudaError(std::string("ncclCommInitRank: ") + c.nccl.GetErrorString((ncclResult_t)r));
        }
        c.rank = rank;
        c.nRanks = nRanks;
        if (nRanks > 1) setupP2P(c);
    });
}

void b200ls_finalize(void) {
    Context& c = ctx();
    if (!c.initialised) return;
    cudaStreamSynchronize(c.stream);
    for (void* q : c.p2p.opened) cudaIpcCloseMemHandle(q);
    c.p2p.opened.clear();
    c.p2p.freeBlocks.clear();
    c.p2p.enabled = false;
    if (c.comm) {
        c.nccl.CommDestroy(c.comm);
        c.comm = nullptr;
    }
}

int b200ls_set_host_comm(int32_t rank, int32_t nRanks, exchange_cb_t exchange, sum_cb_t sum) {
    g_exchangeCb = exchange;
    g_sumCb = sum;
    g_hostRank = rank;
    g_hostRanks = nRanks;
    g_hostCommSet = (exchange != nullptr && sum != nullptr);
    return 0;
}

b200ls_mesh_t b200ls_mesh_create(int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper,
                                 int32_t nInterfaces, const int32_t* ifaceSizes,
                                 const int32_t* const* ifaceFaceCells, const int32_t* ifaceNeighbRank) {
    b200ls_mesh_t mesh = nullptr;
    int rc = guarded([&] {
        std::vector<HostInterface> ifs(nInterfaces);
        for (int i = 0; i < nInterfaces; i++) {
            ifs[i].faceCells.assign(ifaceFaceCells[i], ifaceFaceCells[i] + ifaceSizes[i]);
            if (ifaceNeighbRank[i] >= 0) {
                ifs[i].neighbRank = ifaceNeighbRank[i];
            } else {
                // cyclic half: -(1 + partner patch)
                const int partner = -1 - ifaceNeighbRank[i];
                if (partner >= nInterfaces || partner == i || ifaceNeighbRank[partner] != -1 - i)
                    throw CudaError("cyclic patch " + std::to_string(i) + ": partner patch does not point back");
                if (ifaceSizes[partner] != ifaceSizes[i])
                    throw CudaError("cyclic patch " + std::to_string(i) + ": partner patch has a different size");
                ifs[i].partner = partner;
            }
        }
        std::unique_ptr<b200ls_mesh_s> m(new b200ls_mesh_s);
        m->host.levels.resize(1);
        m->host.nRanks = g_hostCommSet ? g_hostRanks : ctx().nRanks;
        m->host.rank = g_hostCommSet ? g_hostRank : ctx().rank;
        buildLevel(m->host.levels[0], nCells, nFaces, lower, upper, std::move(ifs));
        mesh = m.release();
    });
    return rc == 0 ? mesh : nullptr;
}

void b200ls_mesh_free(b200ls_mesh_t mesh) { delete mesh; }

int b200ls_mesh_n_levels(b200ls_mesh_t mesh) { return mesh ? int(mesh->host.levels.size()) : 0; }

int b200ls_mesh_get_i32(b200ls_mesh_t mesh, int which, int level, const int32_t** data, int64_t* n) {
    return guarded([&] {
        if (!mesh) throw CudaError("null mesh");
        if (level < 0 || level >= int(mesh->host.levels.size())) throw CudaError("level out of range");
        LevelHost& L = mesh->host.levels[level];
        static thread_local std::vector<int32_t> sizes;
        const std::vector<int32_t>* v = nullptr;
        switch (which) {
            case B200LS_LOSORT: v = &L.losort; break;
            case B200LS_OWNER_START: v = &L.ownerStart; break;
            case B200LS_LOSORT_START: v = &L.losortStart; break;
            case B200LS_FWD_LEVEL_OFFSETS: v = &L.fwdOffsets; break;
            case B200LS_FWD_LEVEL_ROWS: v = &L.fwdRows; break;
            case B200LS_BWD_LEVEL_OFFSETS: v = &L.bwdOffsets; break;
            case B200LS_BWD_LEVEL_ROWS: v = &L.bwdRows; break;
            case B200LS_RESTRICT_ADDRESSING: v = &L.restrictAddr; break;
            case B200LS_FACE_RESTRICT_ADDRESSING: v = &L.faceRestrictAddr; break;
            case B200LS_FACE_FLIP_MAP: v = &L.faceFlip; break;
            case B200LS_LOWER_ADDR: v = &L.lower; break;
            case B200LS_UPPER_ADDR: v = &L.upper; break;
            case B200LS_PERM: v = &L.perm; break;
            case B200LS_LPTR: v = &L.Lptr; break;
            case B200LS_LCOL: v = &L.Lcol; break;
            case B200LS_LFACE: v = &L.Lface; break;
            case B200LS_UPTR: v = &L.Uptr; break;
            case B200LS_UCOL: v = &L.Ucol; break;
            case B200LS_UFACE: v = &L.Uface; break;
            case B200LS_FWD_POS: v = &L.fwdPos; break;
            case B200LS_PENCIL_DIMS:
                sizes.clear();
                if (L.pencil.valid)
                    sizes = {L.pencil.nx, L.pencil.ny, L.pencil.nz, L.pencil.WJ, L.pencil.WK, L.pencil.nJ, L.pencil.nK};
                v = &sizes;
                break;
            case B200LS_PENCIL_TILES: {
                static_assert(sizeof(PencilTile) == 10 * sizeof(int32_t), "PencilTile layout");
                *data = reinterpret_cast<const int32_t*>(L.pencil.tiles.data());
                *n = int64_t(L.pencil.tiles.size()) * 10;
                return;
            }
            case B200LS_PENCIL_ORDER: v = &L.pencil.fwdOrder; break;
            case B200LS_LEVEL_SIZES:
                sizes = {L.nCells, L.nFaces};
                v = &sizes;
                break;
            default: throw CudaError("unknown array id");
        }
        if ((which == B200LS_RESTRICT_ADDRESSING || which == B200LS_FACE_RESTRICT_ADDRESSING ||
             which == B200LS_FACE_FLIP_MAP) && !L.hasCoarse) {
            throw CudaError("level has no coarser level");
        }
        *data = v->data();
        *n = int64_t(v->size());
    });
}

int b200ls_mesh_get_iface_i32(b200ls_mesh_t mesh, int which, int level, int iface, const int32_t** data,
                              int64_t* n) {
    return guarded([&] {
        if (!mesh) throw CudaError("null mesh");
        if (level < 0 || level >= int(mesh->host.levels.size())) throw CudaError("level out of range");
        LevelHost& L = mesh->host.levels[level];
        if (iface < 0 || iface >= int(L.interfaces.size())) throw CudaError("interface out of range");
        const std::vector<int32_t>* v = nullptr;
        if (which == B200LS_IFACE_FACE_CELLS) {
            v = &L.interfaces[iface].faceCells;
        } else if (which == B200LS_IFACE_FACE_RESTRICT_ADDRESSING) {
            if (!L.hasCoarse) throw CudaError("level has no coarser level");
            v = &L.patchFaceRestrictAddr[iface];
        } else {
            throw CudaError("unknown interface array id");
        }
        *data = v->data();
        *n = int64_t(v->size());
    });
}

int b200ls_agglomerate(b200ls_mesh_t mesh, const double* faceWeights, int32_t minCellsPerProcessor,
                       int32_t mergeLevels, int32_t forwardStart) {
    int nCoarse = -1;
    int rc = guarded([&] {
        if (!mesh) throw CudaError("null mesh");
        bool fwd = forwardStart < 0 ? g_forward : (forwardStart != 0);
        HostComm hc = makeHostComm();
        nCoarse = agglomerate(mesh->host, faceWeights, minCellsPerProcessor, mergeLevels, fwd,
                              mesh->host.nRanks > 1 ? &hc : nullptr);
        g_forward = fwd;
        mesh->dev.reset();
        mesh->generation++;
    });
    return rc == 0 ? nCoarse : -1;
}

int b200ls_agglomerate_from_maps(b200ls_mesh_t mesh, int32_t nCoarseLevels, const int32_t* const* restrictAddr,
                                 const int32_t* nCoarseCells) {
    int n = -1;
    int rc = guarded([&] {
        if (!mesh) throw CudaError("null mesh");
        HostComm hc = makeHostComm();
        n = agglomerateFromMaps(mesh->host, nCoarseLevels, restrictAddr, nCoarseCells,
                                mesh->host.nRanks > 1 ? &hc : nullptr);
        mesh->dev.reset();
        mesh->generation++;
    });
    return rc == 0 ? n : -1;
}

b200ls_matrix_t b200ls_matrix_create(b200ls_mesh_t mesh) {
    if (!mesh) {
        g_lastError = "null mesh";
        return nullptr;
    }
    b200ls_matrix_t m = new b200ls_matrix_s;
    m->mesh = mesh;
    return m;
}

void b200ls_matrix_free(b200ls_matrix_t m) { delete m; }

int b200ls_matrix_set(b200ls_matrix_t m, const double* diag, const double* upper, const double* lower,
                      const double* const* bou, const double* const* inn) {
    return guarded([&] {
        if (!m) throw CudaError("null matrix");
        matrixSet(m, diag, upper, lower, bou, inn);
    });
}

int b200ls_matrix_set_dev(b200ls_matrix_t m, const double* diag, const double* upper, const double* lower,
                          const double* const* bou, const double* const* inn) {
    return guarded([&] {
        if (!m) throw CudaError("null matrix");
        matrixSet(m, diag, upper, lower, bou, inn, true);
    });
}

// 64-bit fingerprint of host arrays: four independent multiply-rotate lanes over 32-byte blocks (memory-bound)
static inline unsigned long long fpMix(unsigned long long h, unsigned long long v) {
    h ^= v * 0x9E3779B97F4A7C15ull;
    h = (h << 27) | (h >> 37);
    return h * 0x94D049BB133111EBull + 0x2545F4914F6CDD1Dull;
}
static unsigned long long fingerprintOf(const double* data, size_t n, unsigned long long seed) {
    unsigned long long h0 = seed, h1 = seed ^ 0xA5A5A5A5A5A5A5A5ull, h2 = ~seed, h3 = seed + 0x632BE59BD9B4E019ull;
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(data);
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        h0 = fpMix(h0, w[i]);
        h1 = fpMix(h1, w[i + 1]);
        h2 = fpMix(h2, w[i + 2]);
        h3 = fpMix(h3, w[i + 3]);
    }
    for (; i < n; i++) h0 = fpMix(h0, w[i]);
    return fpMix(fpMix(fpMix(fpMix(n, h0), h1), h2), h3);
}

int b200ls_matrix_set_if_changed(b200ls_matrix_t m, const double* diag, const double* upper, const double* lower,
                                 const double* const* bou, const double* const* inn, int32_t* changed) {
    return guarded([&] {
        if (!m) throw CudaError("null matrix");
        const LevelHost& L = m->mesh->host.levels[0];
        unsigned long long h = fingerprintOf(diag, L.nCells, 1);
        h = fpMix(h, fingerprintOf(upper, L.nFaces, 2));
        h = fpMix(h, lower ? fingerprintOf(lower, L.nFaces, 3) : 0x5bd1e995ull);
        for (size_t i = 0; i < L.interfaces.size(); i++) {
            const size_t n = L.interfaces[i].faceCells.size();
            h = fpMix(h, fingerprintOf(bou[i], n, 4 + 2 * i));
            h = fpMix(h, fingerprintOf(inn[i], n, 5 + 2 * i));
        }
        const bool same = m->valuesSet && m->hasFingerprint && m->fingerprint == h &&
                          m->meshGeneration == m->mesh->generation && m->symmetric == (lower == nullptr);
        if (changed) *changed = same ? 0 : 1;
        if (same) return;
        matrixSet(m, diag, upper, lower, bou, inn, false);
        m->fingerprint = h;
        m->hasFingerprint = true;
    });
}

int b200ls_amul(b200ls_matrix_t m, const double* psi, double* Apsi) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        m->stageA.alloc(std::max<size_t>(m->stageA.n, n));
        double* x = m->vec("psi");
        double* y = m->vec("wA");
        h2d(m->stageA.p, psi, n);
        toPositions(m, 0, x, m->stageA.p);
        opAmul(m, 0, y, x);
        toCells(m, 0, m->stageA.p, y);
        d2h(Apsi, m->stageA.p, n);
    });
}

int b200ls_residual(b200ls_matrix_t m, const double* psi, const double* source, double* rA) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        m->stageA.alloc(std::max<size_t>(m->stageA.n, n));
        double* x = m->vec("psi");
        double* b = m->vec("source");
        double* y = m->vec("wA");
        h2d(m->stageA.p, psi, n);
        toPositions(m, 0, x, m->stageA.p);
        h2d(m->stageA.p, source, n);
        toPositions(m, 0, b, m->stageA.p);
        opResidual(m, 0, y, x, b);
        toCells(m, 0, m->stageA.p, y);
        d2h(rA, m->stageA.p, n);
    });
}

int b200ls_sum_a(b200ls_matrix_t m, double* sumA) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        m->stageA.alloc(std::max<size_t>(m->stageA.n, n));
        double* y = m->vec("wA");
        opSumA(m, 0, y);
        toCells(m, 0, m->stageA.p, y);
        d2h(sumA, m->stageA.p, n);
    });
}

int b200ls_precondition(b200ls_matrix_t m, int precond, const double* rA, double* wA) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        m->stageA.alloc(std::max<size_t>(m->stageA.n, n));
        double* x = m->vec("rA");
        double* y = m->vec("wA");
        h2d(m->stageA.p, rA, n);
        toPositions(m, 0, x, m->stageA.p);
        opPrecondition(m, 0, precond, y, x);
        toCells(m, 0, m->stageA.p, y);
        d2h(wA, m->stageA.p, n);
        checkSweepError();
    });
}

int b200ls_reciprocal_d(b200ls_matrix_t m, int precond, double* rD) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        m->stageA.alloc(std::max<size_t>(m->stageA.n, n));
        m->levels[0].rDValid = false;
        ensureFactor(m, 0, precond);
        toCells(m, 0, m->stageA.p, m->levels[0].rD.p);
        d2h(rD, m->stageA.p, n);
        checkSweepError();
    });
}

int b200ls_smooth(b200ls_matrix_t m, int smoother, double* psi, const double* source, int32_t nSweeps) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        m->stageA.alloc(std::max<size_t>(m->stageA.n, n));
        b200ls::Vec& vPsi = m->vecs["psi"];
        b200ls::Vec& vSpare = m->vecs["psiSpare"];
        m->vec("psi");
        m->vec("psiSpare");
        double* b = m->vec("source");
        h2d(m->stageA.p, psi, n);
        toPositions(m, 0, vPsi.buf.p, m->stageA.p);
        h2d(m->stageA.p, source, n);
        toPositions(m, 0, b, m->stageA.p);
        m->levels[0].rDValid = false;
        opSmooth(m, 0, smoother, vPsi.buf.p, vSpare.buf.p, b, nSweeps);
        toCells(m, 0, m->stageA.p, vPsi.buf.p);
        d2h(psi, m->stageA.p, n);
        checkSweepError();
    });
}

void b200ls_controls_default(b200ls_controls* c) {
    memset(c, 0, sizeof(*c));
    c->solver = B200LS_PCG;
    c->precond = B200LS_DIC;
    c->tolerance = 1e-6;
    c->relTol = 0;
    c->maxIter = 1000;
    c->minIter = 0;
    c->nPreSweeps = 0;
    c->preSweepsLevelMultiplier = 1;
    c->maxPreSweeps = 4;
    c->nPostSweeps = 2;
    c->postSweepsLevelMultiplier = 1;
    c->maxPostSweeps = 4;
    c->nFinestSweeps = 2;
    c->scaleCorrection = -1;
    c->nSweeps = 1;
    c->recordHistory = 0;
    c->precSmoother = B200LS_GAUSS_SEIDEL;
    c->nVcycles = 2;
    c->precTolerance = 1e-6;
    c->precRelTol = 0;
}

int b200ls_solve_dev(b200ls_matrix_t m, const b200ls_controls* c, double* psi_dev, const double* source_dev,
                     b200ls_perf* perf) {
    return guarded([&] {
        requireValues(m);
        solveDev(m, *c, psi_dev, source_dev, perf);
    });
}

int b200ls_solve(b200ls_matrix_t m, const b200ls_controls* c, double* psi, const double* source,
                 b200ls_perf* perf) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        m->stageA.alloc(std::max<size_t>(m->stageA.n, n));
        m->stageB.alloc(std::max<size_t>(m->stageB.n, n));
        cudaEvent_t e0, e1, e2, e3;
        B2_CUDA(cudaEventCreate(&e0));
        B2_CUDA(cudaEventCreate(&e1));
        B2_CUDA(cudaEventCreate(&e2));
        B2_CUDA(cudaEventCreate(&e3));
        B2_CUDA(cudaEventRecord(e0, ctx().stream));
        h2d(m->stageA.p, psi, n);
        h2d(m->stageB.p, source, n);
        B2_CUDA(cudaEventRecord(e1, ctx().stream));
        solveDev(m, *c, m->stageA.p, m->stageB.p, perf);
        B2_CUDA(cudaEventRecord(e2, ctx().stream));
        d2h(psi, m->stageA.p, n);
        B2_CUDA(cudaEventRecord(e3, ctx().stream));
        B2_CUDA(cudaEventSynchronize(e3));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e2, e3);
        perf->h2dMs = a + b;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaEventDestroy(e2);
        cudaEventDestroy(e3);
    });
}

int b200ls_time_kernel(b200ls_matrix_t m, int which, int reps, double* ms) {
    return guarded([&] {
        requireValues(m);
        const int n = m->mesh->host.levels[0].nCells;
        double* x = m->vec("psi");
        double* y = m->vec("wA");
        double* r = m->vec("rA");
        cudaStream_t s = ctx().stream;
        // deterministic, finite contents
        if (n) {
            k_fill<<<1024, 256, 0, s>>>(x, 1.0, n);
            k_fill<<<1024, 256, 0, s>>>(r, 1.0, n);
        }
        b200ls::Vec& vPsi = m->vecs["psi"];
        b200ls::Vec& vSpare = m->vecs["psiSpare"];
        m->vec("psiSpare");
        if (which == 1) ensureFactor(m, 0, m->symmetric ? B200LS_DIC : B200LS_DILU);
        cudaEvent_t e0, e1;
        B2_CUDA(cudaEventCreate(&e0));
        B2_CUDA(cudaEventCreate(&e1));
        auto once = [&] {
            switch (which) {
                case 0: opAmul(m, 0, y, x); break;
                case 1: opPrecondition(m, 0, m->symmetric ? B200LS_DIC : B200LS_DILU, y, r); break;
                case 2: opSmooth(m, 0, B200LS_GAUSS_SEIDEL, vPsi.buf.p, vSpare.buf.p, r, 1); break;
                default: throw CudaError("unknown kernel id");
            }
        };
        for (int i = 0; i < 3; i++) once();
        B2_CUDA(cudaEventRecord(e0, s));
        for (int i = 0; i < reps; i++) once();
        B2_CUDA(cudaEventRecord(e1, s));
        B2_CUDA(cudaEventSynchronize(e1));
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        *ms = double(t) / reps;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        checkSweepError();
    });
}

}  // extern "C"
