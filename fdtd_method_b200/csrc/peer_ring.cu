// peer_ring.cu -- z-slab halo exchange by the COPY ENGINES into peer-mapped ghost planes (NVLink 5 / NVSwitch).
//
// Same ring as nccl_ring.cu (the reference's coarray remote GETs of whole planes, coarray/fdtd.F90:90-91, 97-98), but
// no kernel of anybody's runs on an SM: the temporally blocked pass owns every SM (one 512-thread CTA holds the whole
// register file), so NCCL's send/recv kernels have to wait for a CTA to retire before the halo can move, and then
// take that SM away from the pass.  Here every rank PUSHES its boundary planes into the ghost
// planes of its two neighbours with cudaMemcpyAsync (device-to-device over NVLink, DMA engines), and the hand-off is
// four 32-bit sequence flags per rank moved the same way:
//
//     sender (exchange number seq)                               receiver
//     ready(seq) -> both neighbours  ......................  "my previous pass is complete: your pushes may overwrite
//     wait ready(seq) from both neighbours                     my ghost planes, and my boundary planes are final"
//     push the planes (one copy per contiguous plane range)
//     data(seq)  -> both neighbours  ......................  waited for by cuStreamWaitValue32 on the consuming stream,
//                                                             or inside the pass kernel by the CTAs that read ghost
//                                                             planes (fused_kernel_t2.cuh: those chunks are issued last)
//
// Waits are stream memory operations executed by the front end (cuStreamWaitValue32), flag writes are a front-end write
// of a local staging word followed by a 4-byte peer copy ordered after the plane copies on the same stream.
//
// Mappings: one process per GPU -> CUDA IPC handles of the cudaMalloc'ed arrays, exchanged with the two neighbours over
// the NCCL communicator that fdtd_comm_init() has just created (NCCL stays the bootstrap and the fallback transport);
// one process driving several GPUs (fdtd_comm_init_local) -> plain peer access, no NCCL at all.
#include <cuda.h>

#include <cstring>
#include <string>

#include "solver.h"

namespace fdtd_b200 {

typedef CUresult (*stream_value32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

struct PeerRing {
    int up = 0, down = 0;
    int nk_down = 0;                       // planes owned by the lower neighbour: its top ghosts are planes nk_down, nk_down + 1
    // index 0 = lower neighbour (rank - 1), 1 = upper neighbour (rank + 1)
    char* rbase[2][NCOMP][2] = {};         // the neighbour's array bases, mapped here
    unsigned* rflags[2] = {nullptr, nullptr};
    unsigned* lflags = nullptr;            // this rank's flag words (device memory, zeroed)
    unsigned seq = 0;
    bool ipc = false;
    void* opened[2][2 * NCOMP + 1] = {};   // cudaIpcOpenMemHandle results to close (index 0 only when up == down)
    stream_value32_fn wait32 = nullptr, write32 = nullptr;
};

static bool load_mem_ops(PeerRing* r, std::string& err) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        err = "cuStreamWaitValue32 not available";
        return false;
    }
    r->wait32 = reinterpret_cast<stream_value32_fn>(p);
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        err = "cuStreamWriteValue32 not available";
        return false;
    }
    r->write32 = reinterpret_cast<stream_value32_fn>(p);
    return true;
}

static fdtd_status_t alloc_flags(PeerRing* r) {
    FDTD_CUDA_TRY(cudaMalloc(&r->lflags, PEER_FLAG_WORDS * sizeof(unsigned)));
    FDTD_CUDA_TRY(cudaMemset(r->lflags, 0, PEER_FLAG_WORDS * sizeof(unsigned)));
    return FDTD_OK;
}

static void neighbours(const Solver* s, PeerRing* r) {
    const int P = s->cfg.nranks, me = s->cfg.rank;
    r->up = (me + 1) % P;
    r->down = (me + P - 1) % P;
    int b, e;
    fdtd_slab_range_cfg(&s->cfg, r->down, &b, &e);
    r->nk_down = e - b;
}

// ---- one process per GPU: CUDA IPC -----------------------------------------------------------------------------
struct IpcPacket {
    cudaIpcMemHandle_t arr[NCOMP][2];
    cudaIpcMemHandle_t flags;
    int present[NCOMP][2];
    int nk, device;
};

static bool open_packet(PeerRing* r, int side, const IpcPacket& pk, std::string& err) {
    int n = 0;
    for (int c = 0; c < NCOMP; ++c)
        for (int g = 0; g < 2; ++g) {
            if (!pk.present[c][g]) continue;
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, pk.arr[c][g], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); return false; }
            r->rbase[side][c][g] = static_cast<char*>(p);
            r->opened[side][n++] = p;
        }
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, pk.flags, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); err = std::string("cudaIpcOpenMemHandle(flags): ") + cudaGetErrorString(e); return false; }
    r->rflags[side] = static_cast<unsigned*>(p);
    r->opened[side][n++] = p;
    return true;
}

// Called by nccl_init() right after the communicator exists.  `sendrecv` moves `bytes` device bytes to / from a rank
// (one grouped NCCL call), `all_ok` is a min-reduction over the ring: either every rank maps its neighbours or every
// rank stays on the NCCL transport.
fdtd_status_t peer_ring_init_ipc(Solver* s, const PeerBootstrap& boot) {
    if (s->cfg.nranks <= 1 || s->peer) return FDTD_OK;
    PeerRing* r = new PeerRing();
    std::string err;
    int ok = load_mem_ops(r, err) ? 1 : 0;
    neighbours(s, r);
    IpcPacket mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok && alloc_flags(r) != FDTD_OK) { ok = 0; err = fdtd_last_error(); }
    if (ok) {
        for (int c = 0; c < NCOMP && ok; ++c)
            for (int g = 0; g < 2 && ok; ++g) {
                if (!s->base[c][g]) continue;
                if (cudaIpcGetMemHandle(&mine.arr[c][g], s->base[c][g]) != cudaSuccess) { cudaGetLastError(); ok = 0; err = "cudaIpcGetMemHandle failed"; }
                mine.present[c][g] = 1;
            }
        if (ok && cudaIpcGetMemHandle(&mine.flags, r->lflags) != cudaSuccess) { cudaGetLastError(); ok = 0; err = "cudaIpcGetMemHandle(flags) failed"; }
    }
    mine.nk = s->g.nk; mine.device = s->device;
    // packets travel through device buffers (NCCL moves device memory); everybody takes part even after a local failure
    IpcPacket* d = nullptr;   // [0] mine, [1] from down, [2] from up
    FDTD_CUDA_TRY(cudaMalloc(&d, 3 * sizeof(IpcPacket)));
    FDTD_CUDA_TRY(cudaMemcpy(d, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    fdtd_status_t st = boot.exchange(boot.ctx, d, r->up, r->down, d + 1, r->down, d + 2, r->up, sizeof(IpcPacket));
    IpcPacket from[2];
    if (st == FDTD_OK) {
        FDTD_CUDA_TRY(cudaMemcpy(from, d + 1, 2 * sizeof(IpcPacket), cudaMemcpyDeviceToHost));
    } else ok = 0;
    cudaFree(d);
    if (ok) {
        if (!open_packet(r, 0, from[0], err)) ok = 0;
        if (ok && r->up == r->down) {   // two ranks: the same allocations serve both directions
            std::memcpy(r->rbase[1], r->rbase[0], sizeof(r->rbase[0]));
            r->rflags[1] = r->rflags[0];
        } else if (ok && !open_packet(r, 1, from[1], err)) ok = 0;
    }
    int all = ok;
    if (st == FDTD_OK) st = boot.all_min(boot.ctx, &all);
    if (st != FDTD_OK || !all) {
        // stay on NCCL (all ranks agree); not an error
        for (int side = 0; side < 2; ++side)
            for (void* p : r->opened[side]) if (p) cudaIpcCloseMemHandle(p);
        if (r->lflags) cudaFree(r->lflags);
        delete r;
        cudaGetLastError();
        s->peer_note = ok ? "a peer could not map this rank's memory" : err;
        return st;
    }
    r->ipc = true;
    s->peer = r;
    return FDTD_OK;
}

// ---- one process, several GPUs: plain peer access ----------------------------------------------------------------
fdtd_status_t peer_ring_init_local(Solver** all, int n) {
    for (int i = 0; i < n; ++i) {
        Solver* s = all[i];
        if (!s || s->cfg.nranks != n || s->cfg.rank != i) return fail(FDTD_ERR_BAD_ARGUMENT, "comm_init_local: solvers[r] must be rank r of n");
        if (s->peer || s->comm) return fail(FDTD_ERR_STATE, "communicator already initialised");
    }
    if (n <= 1) return FDTD_OK;
    for (int i = 0; i < n; ++i) {
        Solver* s = all[i];
        FDTD_CUDA_TRY(cudaSetDevice(s->device));
        PeerRing* r = new PeerRing();
        std::string err;
        if (!load_mem_ops(r, err)) { delete r; return fail(FDTD_ERR_CUDA, err); }
        neighbours(s, r);
        fdtd_status_t st = alloc_flags(r);
        if (st != FDTD_OK) { delete r; return st; }
        for (int nb : {r->up, r->down}) {
            const int pd = all[nb]->device;
            if (pd == s->device) continue;
            int can = 0;
            FDTD_CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->device, pd));
            if (!can) { cudaFree(r->lflags); delete r; return fail(FDTD_ERR_CUDA, "comm_init_local: no peer access between the devices of neighbouring slabs"); }
            cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaFree(r->lflags); delete r; return cuda_fail(e, "cudaDeviceEnablePeerAccess"); }
            cudaGetLastError();
        }
        s->peer = r;
    }
    for (int i = 0; i < n; ++i) {
        PeerRing* r = all[i]->peer;
        const int nb[2] = {r->down, r->up};
        for (int side = 0; side < 2; ++side) {
            const Solver* o = all[nb[side]];
            for (int c = 0; c < NCOMP; ++c)
                for (int g = 0; g < 2; ++g) r->rbase[side][c][g] = static_cast<char*>(o->base[c][g]);
            r->rflags[side] = o->peer->lflags;
        }
    }
    return FDTD_OK;
}

void peer_ring_destroy(Solver* s) {
    PeerRing* r = s->peer;
    if (!r) return;
    if (r->ipc)
        for (int side = 0; side < 2; ++side)
            for (void* p : r->opened[side]) if (p) cudaIpcCloseMemHandle(p);
    if (r->lflags) cudaFree(r->lflags);
    delete r;
    s->peer = nullptr;
}

// Wait on `q` until both neighbours' planes of the latest exchange have landed in this rank's ghost planes.
fdtd_status_t peer_wait_data(Solver* s, cudaStream_t q) {
    PeerRing* r = s->peer;
    if (!r || !r->seq) return FDTD_OK;
    CUresult a = r->wait32(q, (CUdeviceptr)(r->lflags + PEER_F_DATA_FROM_DOWN), r->seq, CU_STREAM_WAIT_VALUE_GEQ);
    CUresult b = r->wait32(q, (CUdeviceptr)(r->lflags + PEER_F_DATA_FROM_UP), r->seq, CU_STREAM_WAIT_VALUE_GEQ);
    if (a != CUDA_SUCCESS || b != CUDA_SUCCESS) return fail(FDTD_ERR_CUDA, "cuStreamWaitValue32 failed");
    return FDTD_OK;
}

const unsigned* peer_data_flags(const Solver* s) { return s->peer ? s->peer->lflags + PEER_F_DATA_FROM_DOWN : nullptr; }
unsigned* peer_error_word(const Solver* s) { return s->peer ? s->peer->lflags + PEER_F_ERR : nullptr; }
unsigned peer_last_seq(const Solver* s) { return s->peer ? s->peer->seq : 0u; }

#define FDTD_CU_TRY(expr)                                                                                       \
    do {                                                                                                        \
        CUresult _r = (expr);                                                                                   \
        if (_r != CUDA_SUCCESS) return fail(FDTD_ERR_CUDA, std::string("driver error ") + std::to_string((int)_r) + " in " #expr); \
    } while (0)

// Push exchange.  x[i].to_upper selects the neighbour, x[i].dst_plane is the plane index in the RECEIVER's numbering
// (-2, -1 = its bottom ghosts; 0, 1 here mean its top ghosts nk_peer, nk_peer + 1 -- the sender adds the peer's plane
// count).  wait_data: also wait (on `q`) until both neighbours' planes have landed here; false when the consumer
// waits itself (the T2 pass kernel).  ev_start / ev_end (optional, timing events) bracket the plane copies.
fdtd_status_t peer_exchange(Solver* s, const PlaneXfer* x, int n, cudaStream_t q, bool wait_data, cudaEvent_t ev_start, cudaEvent_t ev_end) {
    PeerRing* r = s->peer;
    const unsigned seq = ++r->seq;
    unsigned* stage = r->lflags + PEER_F_STAGE + (seq % PEER_STAGE_SLOTS);
    const size_t pb = (size_t)s->g.plane * s->esz;
    FDTD_CU_TRY(r->write32(q, (CUdeviceptr)stage, seq, 0));
    // I am the upper neighbour of my lower neighbour, and the other way round
    FDTD_CUDA_TRY(cudaMemcpyAsync(r->rflags[0] + PEER_F_READY_FROM_UP, stage, 4, cudaMemcpyDeviceToDevice, q));
    FDTD_CUDA_TRY(cudaMemcpyAsync(r->rflags[1] + PEER_F_READY_FROM_DOWN, stage, 4, cudaMemcpyDeviceToDevice, q));
    FDTD_CU_TRY(r->wait32(q, (CUdeviceptr)(r->lflags + PEER_F_READY_FROM_DOWN), seq, CU_STREAM_WAIT_VALUE_GEQ));
    FDTD_CU_TRY(r->wait32(q, (CUdeviceptr)(r->lflags + PEER_F_READY_FROM_UP), seq, CU_STREAM_WAIT_VALUE_GEQ));
    if (ev_start) FDTD_CUDA_TRY(cudaEventRecord(ev_start, q));
    for (int i = 0; i < n; ++i) {
        const int side = x[i].to_upper ? 1 : 0;
        const long long plane = x[i].to_upper ? x[i].dst_plane : (long long)r->nk_down + x[i].dst_plane;
        char* dst = r->rbase[side][x[i].comp][x[i].gen];
        if (!dst) return fail(FDTD_ERR_STATE, "peer exchange: the neighbour has no such array");
        dst += (size_t)(GHOST_PLANES + plane) * pb;
        FDTD_CUDA_TRY(cudaMemcpyAsync(dst, x[i].send, x[i].bytes, cudaMemcpyDeviceToDevice, q));
    }
    FDTD_CUDA_TRY(cudaMemcpyAsync(r->rflags[0] + PEER_F_DATA_FROM_UP, stage, 4, cudaMemcpyDeviceToDevice, q));
    FDTD_CUDA_TRY(cudaMemcpyAsync(r->rflags[1] + PEER_F_DATA_FROM_DOWN, stage, 4, cudaMemcpyDeviceToDevice, q));
    if (ev_end) FDTD_CUDA_TRY(cudaEventRecord(ev_end, q));
    if (wait_data) {
        FDTD_CU_TRY(r->wait32(q, (CUdeviceptr)(r->lflags + PEER_F_DATA_FROM_DOWN), seq, CU_STREAM_WAIT_VALUE_GEQ));
        FDTD_CU_TRY(r->wait32(q, (CUdeviceptr)(r->lflags + PEER_F_DATA_FROM_UP), seq, CU_STREAM_WAIT_VALUE_GEQ));
    }
    return FDTD_OK;
}

}  // namespace fdtd_b200
