// Multi-GPU collection of the correspondence rows over peer memory (NVLink / NVSwitch), SURVEY.md section 8e step (4):
// "gather(v) of (pair id, K, K x 6 f64) to rank 0, concatenated in (i,j) pair order" -- the order Frame::corres_kps is
// filled in by the i<j loop of src/diasss2.cpp:88-97 and FEAmatcher.cpp:35-45.
//
// One process per GPU.  Every rank owns a small block of device memory (cudaMalloc) that it exports with a CUDA IPC
// handle; the ranks exchange the 64-byte handles once (torch.distributed) and map each other's blocks.  Per step:
//   rank r  scan kernel    publishes its row total K_r into slot [r] of every rank's totals array       (8-byte stores)
//   rank r  emit kernel    waits for the totals of ranks 0..r-1 in ITS OWN totals array, then writes its rows and
//                          per-pair counts straight into rank 0's block at row offset K_0 + .. + K_{r-1}   (NVLink stores)
//   rank r  done kernel    system-scope fence, then raises done[r] in rank 0's block
//   rank 0  wait kernel    spins until done[0..W-1] carry the step's sequence number; a scan turns the counts into offsets
// No host synchronisation, no padding, no reorder: pair blocks are contiguous in pair order (shard.Plan), so the
// concatenation IS the (i,j) order.  Everything is double-buffered on the parity of the sequence number, so step s+1 may
// be written while step s is still being read on rank 0.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "dsx_internal.cuh"

struct dsx_peer {
    int rank = 0, world = 1, n_pairs = 0;
    long long cap_rows = 0;                // rows per parity half (rank 0's block only)
    int device = 0;
    uint8_t* local = nullptr;
    uint8_t* remote[dsx::kMaxPeers] = {nullptr};
    bool opened[dsx::kMaxPeers] = {false};
    int32_t* l_cnt = nullptr;              // this rank's per-pair counts / offsets of the step in flight (local scratch)
    int32_t* l_off = nullptr;
    size_t o_totals = 0, o_done = 0, o_cnt = 0, o_off = 0, o_rows = 0, bytes = 0;
    unsigned long long timeout_ns = dsx::kPeerTimeoutNs;   // DSX_PEER_TIMEOUT_MS
};

namespace {
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// identical on every rank (only the rows region differs: rank 0 alone has one)
void layout(dsx_peer* p) {
    size_t o = 0;
    p->o_totals = o; o += align_up(sizeof(unsigned long long) * 2 * p->world, 256);
    p->o_done = o;   o += align_up(sizeof(unsigned) * 2 * p->world, 256);
    p->o_cnt = o;    o += align_up(sizeof(int32_t) * 2 * (size_t)std::max(p->n_pairs, 1), 256);
    p->o_off = o;    o += align_up(sizeof(int32_t) * 2 * ((size_t)p->n_pairs + 1), 256);
    p->o_rows = o;
    if (p->rank == 0) o += sizeof(double) * 6 * 2 * (size_t)p->cap_rows;
    p->bytes = o;
}
}  // namespace

extern "C" {

int dsx_peer_create(dsx_ctx* ctx, int rank, int world, int n_pairs_total, int64_t cap_rows, dsx_peer** out, uint8_t* handle) {
    using namespace dsx;
    if (!ctx || !out || !handle || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_pairs_total < 0 || cap_rows < 0) {
        set_error("dsx_peer_create: bad argument");
        return DSX_ERR_INVALID;
    }
    DSX_CUDA(cudaSetDevice(ctx->device));
    dsx_peer* p = new dsx_peer();
    p->rank = rank; p->world = world; p->n_pairs = n_pairs_total; p->cap_rows = cap_rows; p->device = ctx->device;
    if (const char* t = getenv("DSX_PEER_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) p->timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    layout(p);
    cudaError_t e = cudaMalloc((void**)&p->local, p->bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->l_cnt, sizeof(int32_t) * (size_t)std::max(n_pairs_total, 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->l_off, sizeof(int32_t) * ((size_t)n_pairs_total + 1));
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->o_rows);            // sequence numbers start at 1
    if (e == cudaSuccess) e = cudaDeviceSynchronize();                       // (peers may write as soon as they hold the handle)
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) {
        set_error(std::string("dsx_peer_create: ") + cudaGetErrorString(e));
        dsx_peer_destroy(p);
        return e == cudaErrorMemoryAllocation ? DSX_ERR_NOMEM : DSX_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == DSX_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    memcpy(handle, &h, sizeof(h));
    p->remote[rank] = p->local;
    *out = p;
    return DSX_OK;
}

int dsx_peer_connect(dsx_peer* p, const uint8_t* handles) {
    using namespace dsx;
    if (!p || !handles) { set_error("dsx_peer_connect: null argument"); return DSX_ERR_INVALID; }
    DSX_CUDA(cudaSetDevice(p->device));
    for (int q = 0; q < p->world; q++) {
        if (q == p->rank || p->opened[q]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * DSX_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error(std::string("dsx_peer_connect: cudaIpcOpenMemHandle(rank ") + std::to_string(q) + "): " + cudaGetErrorString(e));
            cudaGetLastError();
            return DSX_ERR_CUDA;
        }
        p->remote[q] = (uint8_t*)ptr;
        p->opened[q] = true;
    }
    return DSX_OK;
}

int dsx_peer_connect_local(dsx_peer* p, int q, dsx_peer* other) {
    using namespace dsx;
    if (!p || !other || q < 0 || q >= p->world || other->rank != q || other->world != p->world || other->n_pairs != p->n_pairs ||
        other->cap_rows != p->cap_rows) {
        set_error("dsx_peer_connect_local: bad argument");
        return DSX_ERR_INVALID;
    }
    if (other->device != p->device) {
        DSX_CUDA(cudaSetDevice(p->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            set_error(std::string("dsx_peer_connect_local: ") + cudaGetErrorString(e));
            return DSX_ERR_CUDA;
        }
        cudaGetLastError();
    }
    p->remote[q] = other->local;
    return DSX_OK;
}

int dsx_match_pairs_peer(dsx_ctx* ctx, dsx_peer* p, const dsx_features_dev* feats, const int32_t* img_id, const int32_t* img_rows,
                         const double* bbox, const int32_t* pairs, int n_pairs, int pair_begin, int seq) {
    using namespace dsx;
    if (!ctx || !p || !feats || !img_id || !img_rows || !bbox || (n_pairs > 0 && !pairs) || seq < 1 || n_pairs < 0 || pair_begin < 0 ||
        pair_begin + n_pairs > p->n_pairs) {
        set_error("dsx_match_pairs_peer: bad argument");
        return DSX_ERR_INVALID;
    }
    for (int q = 0; q < p->world; q++)
        if (!p->remote[q]) { set_error("dsx_match_pairs_peer: peers not connected"); return DSX_ERR_INVALID; }
    for (int i = 0; i < 2 * n_pairs; i++)
        if (pairs[i] < 0 || pairs[i] >= feats->n_images) { set_error("pair index out of range"); return DSX_ERR_INVALID; }
    const int par = seq & 1;
    if (n_pairs > 0) {
        DSX_TRY(match_begin(ctx, feats, img_id, img_rows, bbox, pairs, nullptr, n_pairs, nullptr, nullptr, nullptr));
        DSX_TRY(match_stage(ctx, feats, 0, feats->n_images, 0, n_pairs));
    } else {
        ctx->mplan.n_pairs = 0; ctx->mplan.has_slots = false;
    }
    PeerPub pub; memset(&pub, 0, sizeof(pub));
    pub.world = p->world; pub.rank = p->rank; pub.seq = (unsigned)seq;
    for (int q = 0; q < p->world; q++)
        pub.totals[q] = reinterpret_cast<unsigned long long*>(p->remote[q] + p->o_totals) + (size_t)par * p->world;
    PeerSink sink; memset(&sink, 0, sizeof(sink));
    uint8_t* r0 = p->remote[0];
    // (every rank is created with the same n_pairs_total and cap_rows; only rank 0's block has the rows region)
    sink.rows6 = reinterpret_cast<double*>(r0 + p->o_rows) + (size_t)par * 6 * (size_t)p->cap_rows;
    sink.cap_rows = p->cap_rows;
    sink.cnt_dst = reinterpret_cast<int32_t*>(r0 + p->o_cnt) + (size_t)par * std::max(p->n_pairs, 1) + pair_begin;
    sink.my_totals = reinterpret_cast<const unsigned long long*>(p->local + p->o_totals) + (size_t)par * p->world;
    sink.rank = p->rank; sink.seq = (unsigned)seq; sink.timeout_ns = p->timeout_ns;
    unsigned* done_slot = reinterpret_cast<unsigned*>(r0 + p->o_done) + (size_t)par * p->world + p->rank;
    return match_finish_peer(ctx, feats, p->l_cnt, p->l_off, pub, sink, done_slot);
}

int dsx_peer_collect(dsx_ctx* ctx, dsx_peer* p, int seq, const int32_t** corr_count, const int32_t** corr_offset, const double** rows6) {
    using namespace dsx;
    if (!ctx || !p || seq < 1 || p->rank != 0) { set_error("dsx_peer_collect: rank 0 only"); return DSX_ERR_INVALID; }
    const int par = seq & 1;
    const unsigned* done = reinterpret_cast<const unsigned*>(p->local + p->o_done) + (size_t)par * p->world;
    int32_t* cnt = reinterpret_cast<int32_t*>(p->local + p->o_cnt) + (size_t)par * std::max(p->n_pairs, 1);
    int32_t* off = reinterpret_cast<int32_t*>(p->local + p->o_off) + (size_t)par * ((size_t)p->n_pairs + 1);
    DSX_TRY(peer_wait_and_scan(ctx, done, p->world, (unsigned)seq, cnt, off, p->n_pairs, p->timeout_ns));
    if (corr_count) *corr_count = cnt;
    if (corr_offset) *corr_offset = off;
    if (rows6) *rows6 = reinterpret_cast<const double*>(p->local + p->o_rows) + (size_t)par * 6 * (size_t)p->cap_rows;
    return DSX_OK;
}

void dsx_peer_destroy(dsx_peer* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (int q = 0; q < p->world; q++)
        if (p->opened[q] && p->remote[q]) cudaIpcCloseMemHandle(p->remote[q]);
    if (p->local) cudaFree(p->local);
    if (p->l_cnt) cudaFree(p->l_cnt);
    if (p->l_off) cudaFree(p->l_off);
    delete p;
}

}  // extern "C"
