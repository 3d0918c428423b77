// nccl_ring.cu -- z-slab ring halo exchange over NCCL send/recv (NVLink 5 / NVSwitch).
//
// B200 equivalent of the reference's only distributed code, the coarray remote GETs of whole planes
// in coarray/fdtd.F90:90-91 (Bx,By top plane from the predecessor image) and :97-98 (Ex,Ey bottom
// plane from the successor image).  k is the slowest axis, so a plane of one component is one
// contiguous pitch*Nj*sizeof(T) block and an exchange is one ncclSend/ncclRecv pair per plane inside a
// single ncclGroup.  libnccl is dlopen'ed lazily: a single-GPU solver never touches it, and inside a
// torch process the already-loaded torch-bundled libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "solver.h"

namespace fdtd_b200 {

typedef struct { char internal[FDTD_NCCL_UNIQUE_ID_BYTES]; } nccl_uid_t;
typedef int nccl_result_t;
enum { NCCL_CHAR = 0, NCCL_INT32 = 2, NCCL_MIN = 3 };   // ncclInt8 / ncclChar, ncclInt32, ncclMin

struct NcclApi {
    void* lib = nullptr;
    nccl_result_t (*GetUniqueId)(nccl_uid_t*) = nullptr;
    nccl_result_t (*CommInitRank)(void**, int, nccl_uid_t, int) = nullptr;
    nccl_result_t (*CommDestroy)(void*) = nullptr;
    nccl_result_t (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    nccl_result_t (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    nccl_result_t (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    nccl_result_t (*GroupStart)() = nullptr;
    nccl_result_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(nccl_result_t) = nullptr;
};

static NcclApi* load_nccl(std::string& err) {
    static NcclApi api;
    if (api.lib) return &api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return nullptr; }
#define LOAD(field, sym)                                                            \
    *(void**)(&api.field) = dlsym(api.lib, sym);                                    \
    if (!api.field) { err = std::string("libnccl lacks ") + sym; api.lib = nullptr; return nullptr; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return &api;
}

static fdtd_status_t nccl_fail(NcclApi* api, nccl_result_t r, const char* what) {
    return fail(FDTD_ERR_NCCL, std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(r) : "?"));
}

fdtd_status_t nccl_unique_id(void* out, size_t cap) {
    if (!out || cap < FDTD_NCCL_UNIQUE_ID_BYTES) return fail(FDTD_ERR_BAD_ARGUMENT, "unique id buffer too small");
    std::string err;
    NcclApi* api = load_nccl(err);
    if (!api) return fail(FDTD_ERR_NCCL, err);
    nccl_uid_t id;
    nccl_result_t r = api->GetUniqueId(&id);
    if (r != 0) return nccl_fail(api, r, "ncclGetUniqueId");
    std::memcpy(out, &id, sizeof(id));
    return FDTD_OK;
}

fdtd_status_t nccl_init(Solver* s, const void* id, size_t bytes) {
    if (s->cfg.nranks <= 1) return FDTD_OK;
    if (!id || bytes < FDTD_NCCL_UNIQUE_ID_BYTES) return fail(FDTD_ERR_BAD_ARGUMENT, "bad NCCL unique id");
    if (s->comm) return fail(FDTD_ERR_STATE, "communicator already initialised");
    std::string err;
    NcclApi* api = load_nccl(err);
    if (!api) return fail(FDTD_ERR_NCCL, err);
    nccl_uid_t uid;
    std::memcpy(&uid, id, sizeof(uid));
    FDTD_CUDA_TRY(cudaSetDevice(s->device));
    nccl_result_t r = api->CommInitRank(&s->comm, s->cfg.nranks, uid, s->cfg.rank);
    if (r != 0) return nccl_fail(api, r, "ncclCommInitRank");
    s->nccl = api;
    // Default transport: copy engines into peer-mapped ghost planes (peer_ring.cu); the communicator carries the CUDA IPC
    // handles to the neighbours and stays as the fallback when the ring cannot be mapped (FDTD_B200_TRANSPORT=nccl
    // forces it -- every rank must use the same setting).
    const char* tr = std::getenv("FDTD_B200_TRANSPORT");
    if (tr && std::strcmp(tr, "nccl") == 0) { s->peer_note = "FDTD_B200_TRANSPORT=nccl"; return FDTD_OK; }
    PeerBootstrap boot;
    boot.ctx = s;
    boot.exchange = [](void* ctx, const void* send, int up, int down, void* from_down, int down2, void* from_up, int up2, size_t bytes) -> fdtd_status_t {
        Solver* s = static_cast<Solver*>(ctx);
        NcclApi* api = s->nccl;
        nccl_result_t r = api->GroupStart();
        if (r == 0) r = api->Send(send, bytes, NCCL_CHAR, up, s->comm, s->stream);
        if (r == 0) r = api->Send(send, bytes, NCCL_CHAR, down, s->comm, s->stream);
        if (r == 0) r = api->Recv(from_down, bytes, NCCL_CHAR, down2, s->comm, s->stream);
        if (r == 0) r = api->Recv(from_up, bytes, NCCL_CHAR, up2, s->comm, s->stream);
        if (r != 0) { api->GroupEnd(); return nccl_fail(api, r, "ncclSend/Recv (IPC handles)"); }
        r = api->GroupEnd();
        if (r != 0) return nccl_fail(api, r, "ncclGroupEnd (IPC handles)");
        FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
        return FDTD_OK;
    };
    boot.all_min = [](void* ctx, int* value) -> fdtd_status_t {
        Solver* s = static_cast<Solver*>(ctx);
        int* d = nullptr;
        FDTD_CUDA_TRY(cudaMalloc(&d, sizeof(int)));
        FDTD_CUDA_TRY(cudaMemcpy(d, value, sizeof(int), cudaMemcpyHostToDevice));
        nccl_result_t r = s->nccl->AllReduce(d, d, 1, NCCL_INT32, NCCL_MIN, s->comm, s->stream);
        if (r != 0) { cudaFree(d); return nccl_fail(s->nccl, r, "ncclAllReduce"); }
        FDTD_CUDA_TRY(cudaStreamSynchronize(s->stream));
        FDTD_CUDA_TRY(cudaMemcpy(value, d, sizeof(int), cudaMemcpyDeviceToHost));
        cudaFree(d);
        return FDTD_OK;
    };
    return peer_ring_init_ipc(s, boot);
}

void nccl_destroy(Solver* s) {
    peer_ring_destroy(s);
    if (s->comm && s->nccl) s->nccl->CommDestroy(s->comm);
    s->comm = nullptr;
}

fdtd_status_t nccl_exchange(Solver* s, const PlaneXfer* x, int n, cudaStream_t stream) {
    if (!s->comm) return fail(FDTD_ERR_STATE, "multi-rank solver used before fdtd_comm_init()");
    NcclApi* api = s->nccl;
    nccl_result_t r = api->GroupStart();
    if (r != 0) return nccl_fail(api, r, "ncclGroupStart");
    for (int i = 0; i < n; ++i) {
        r = api->Send(x[i].send, x[i].bytes, NCCL_CHAR, x[i].peer_send, s->comm, stream);
        if (r != 0) return nccl_fail(api, r, "ncclSend");
        r = api->Recv(x[i].recv, x[i].bytes, NCCL_CHAR, x[i].peer_recv, s->comm, stream);
        if (r != 0) return nccl_fail(api, r, "ncclRecv");
    }
    r = api->GroupEnd();
    if (r != 0) return nccl_fail(api, r, "ncclGroupEnd");
    return FDTD_OK;
}

}  // namespace fdtd_b200
