// NCCL communicator owned by libsrt.so: the one real exchange step of the path, the film reduce /
// gather between the GPUs of a node (BASELINE.json north_star; the reference is single-GPU and has
// no counterpart).  libnccl.so.2 is opened on first use, so single-GPU users of the
// library do not need NCCL at all; nothing here falls back to a host path.
#include "srt_host.hpp"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <mutex>

namespace srt {

namespace {
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclReduceScatter) ReduceScatter = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
};
std::mutex g_api_mu;
NcclApi g_api;
bool g_api_ok = false;

const NcclApi* api() {
    std::lock_guard<std::mutex> lock(g_api_mu);
    if (g_api_ok) return &g_api;
    // the soname: a process that already carries an NCCL (e.g. the one bundled with PyTorch) shares it
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) { set_error(std::string("libsrt: cannot load libnccl.so.2: ") + dlerror()); return nullptr; }
    g_api.handle = h;
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) ok = false; return p; };
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
    g_api.GetVersion = (decltype(g_api.GetVersion))sym("ncclGetVersion");
    g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
    g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
    g_api.ReduceScatter = (decltype(g_api.ReduceScatter))sym("ncclReduceScatter");
    g_api.AllGather = (decltype(g_api.AllGather))sym("ncclAllGather");
    g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
    g_api.Send = (decltype(g_api.Send))sym("ncclSend");
    g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
    if (!ok) { set_error("libsrt: libnccl.so.2 lacks a required entry point"); return nullptr; }
    g_api_ok = true;
    return &g_api;
}
bool nccl_ok(const NcclApi* A, ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    set_error(std::string("NCCL error in ") + what + ": " + A->GetErrorString(r));
    return false;
}
}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;  // for the small control collectives (barrier, max)
    void* scratch = nullptr;        // 64 bytes of device memory for them
};

#define SRT_NCCL(call)                                    \
    do {                                                  \
        if (!nccl_ok(A, (call), #call)) return false;     \
    } while (0)

bool comm_unique_id(unsigned char id[SRT_NCCL_UNIQUE_ID_BYTES]) {
    static_assert(SRT_NCCL_UNIQUE_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");
    const NcclApi* A = api();
    if (!A) return false;
    ncclUniqueId u;
    SRT_NCCL(A->GetUniqueId(&u));
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return true;
}

Comm* comm_create(const unsigned char id[SRT_NCCL_UNIQUE_ID_BYTES], int rank, int world) {
    const NcclApi* A = api();
    if (!A) return nullptr;
    if (world < 1 || rank < 0 || rank >= world) { set_error("srt_comm_create: bad rank / world"); return nullptr; }
    auto* c = new Comm();
    c->rank = rank; c->world = world;
    ncclUniqueId u;
    memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    if (cudaGetDevice(&c->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c->scratch, 64) != cudaSuccess) {
        set_error("srt_comm_create: CUDA set-up failed");
        comm_destroy(c);
        return nullptr;
    }
    if (!nccl_ok(A, A->CommInitRank(&c->comm, world, u, rank), "ncclCommInitRank")) { c->comm = nullptr; comm_destroy(c); return nullptr; }
    return c;
}
void comm_destroy(Comm* c) {
    if (!c) return;
    if (c->comm && g_api_ok) g_api.CommDestroy(c->comm);
    if (c->scratch) cudaFree(c->scratch);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
int comm_rank(const Comm* c) { return c->rank; }
int comm_world(const Comm* c) { return c->world; }
int comm_device(const Comm* c) { return c->device; }
int comm_nccl_version() {
    const NcclApi* A = api();
    int v = 0;
    if (A) A->GetVersion(&v);
    return v;
}

// sum of the three XYZ planes over all ranks, scattered: afterwards rank r holds elements [r*cnt, (r+1)*cnt) of every plane
bool comm_film_reduce_scatter(Comm* c, float* acc, size_t plane_stride, size_t cnt, cudaStream_t st) {
    const NcclApi* A = api();
    if (!A) return false;
    SRT_NCCL(A->GroupStart());
    for (int p = 0; p < 3; p++) {
        float* plane = acc + (size_t)p * plane_stride;
        SRT_NCCL(A->ReduceScatter(plane, plane + (size_t)c->rank * cnt, cnt, ncclFloat, ncclSum, c->comm, st));
    }
    SRT_NCCL(A->GroupEnd());
    return true;
}
// the inverse hand-out: every rank ends up with all slices of all three planes
bool comm_film_all_gather(Comm* c, float* acc, size_t plane_stride, size_t cnt, cudaStream_t st) {
    const NcclApi* A = api();
    if (!A) return false;
    SRT_NCCL(A->GroupStart());
    for (int p = 0; p < 3; p++) {
        float* plane = acc + (size_t)p * plane_stride;
        SRT_NCCL(A->AllGather(plane + (size_t)c->rank * cnt, plane, cnt, ncclFloat, c->comm, st));
    }
    SRT_NCCL(A->GroupEnd());
    return true;
}
// `bytes` bytes from every rank to rank 0, which receives them rank-major into recv_all
bool comm_gather_bytes(Comm* c, const unsigned char* send, unsigned char* recv_all, size_t bytes, cudaStream_t st) {
    const NcclApi* A = api();
    if (!A) return false;
    if (c->world == 1) {
        if (send != recv_all && cudaMemcpyAsync(recv_all, send, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { set_error("gather copy failed"); return false; }
        return true;
    }
    SRT_NCCL(A->GroupStart());
    if (c->rank == 0) {
        for (int r = 1; r < c->world; r++) SRT_NCCL(A->Recv(recv_all + (size_t)r * bytes, bytes, ncclUint8, r, c->comm, st));
    } else {
        SRT_NCCL(A->Send(send, bytes, ncclUint8, 0, c->comm, st));
    }
    SRT_NCCL(A->GroupEnd());
    if (c->rank == 0 && send != recv_all && cudaMemcpyAsync(recv_all, send, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { set_error("gather copy failed"); return false; }
    return true;
}
bool comm_all_reduce_u64_sum(Comm* c, unsigned long long* dev, size_t n, cudaStream_t st) {
    const NcclApi* A = api();
    if (!A) return false;
    SRT_NCCL(A->AllReduce(dev, dev, n, ncclUint64, ncclSum, c->comm, st));
    return true;
}
// max over ranks of one host double (device-timed milliseconds in bench.py); doubles as a barrier
bool comm_max_double(Comm* c, double* v) {
    const NcclApi* A = api();
    if (!A) return false;
    if (cudaMemcpyAsync(c->scratch, v, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { set_error("comm_max: copy failed"); return false; }
    SRT_NCCL(A->AllReduce(c->scratch, c->scratch, 1, ncclDouble, ncclMax, c->comm, c->stream));
    if (cudaMemcpyAsync(v, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) { set_error("comm_max: read back failed"); return false; }
    return true;
}

}  // namespace srt
