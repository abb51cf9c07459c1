// K2 -- FAST-9/16 detection per ~30 px cell with in-cell 3x3 non-max suppression and the 12 -> 7
// per-cell threshold fallback.  Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree
// (ORBextractor.cpp:771-829) and the cv::FAST(roi, kps, th, true) calls inside it (:809, :814;
// OpenCV features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>).
//
// Semantics reproduced:
//   * m(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min d_k, min -d_k), d_k = I(p)-I(ring k);
//     corner at threshold t <=> m > t; score = m-1 (independent of t).
//   * scores exist only for ROI pixels x in [3,w-3), y in [3,h-3); everything else counts as 0, so NMS never
//     crosses a cell: the per-cell score tile below has a zero border.
//   * keypoint <=> score >= t and score > all 8 neighbours; a cell falls back to minThFAST iff it has no
//     keypoint at iniThFAST.  NMS against the score map of the lower threshold is identical (sub-threshold
//     neighbours are smaller than any score >= t).
//   * output order inside a cell is row-major; coordinates are relative to the ROI origin.
//
// Mapping: one CTA = 8 warps = 8 horizontally adjacent cells of one cell row; the shared (hCell+6) x
// (8*wCell+6) pixel strip is staged in shared memory once with 32-bit coalesced loads.  Each warp then works
// on its own cell, warp-synchronously:
//   pass 1  SIMD-in-word rejection test on 4 pixels per lane (VABSDIFF4 + carry-free byte compares) on the
//           four opposite ring pairs (0,8),(4,12),(2,10),(6,14): a 9-arc contains one pixel of every opposite
//           pair, so a pixel where both members of some pair are within t of the centre cannot be a corner.
//           Survivors are appended, in row-major order, to a shared-memory ring queue by ballot compaction.
//   pass 2  whenever 32 survivors are queued: full 16-pixel arc measure with 3-input min/max (VIMNMX3),
//           score written to the cell's score tile, corners appended (ordered) to the corner list.
//   pass 3  NMS over the corner list (8 neighbour reads from the tile), threshold decision by ballot,
//           ordered emission into the cell's staging slot + count.
// Bound: integer issue rate; HBM traffic is one read of every level.
#include "dsx_internal.cuh"

namespace dsx {

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ uint32_t ld4(const uint8_t* s, int off) {
    // unaligned 4-byte window from shared memory: two aligned words + funnel shift
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (off & ~3));
    return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}

// per-byte (x > t) -> 0x80 in that byte; carry-free
__device__ __forceinline__ uint32_t gt_bytes(uint32_t x, uint32_t k7) {
    return (((x & 0x7f7f7f7fu) + k7) | x) & 0x80808080u;
}

__device__ __forceinline__ int arc_measure(const uint8_t* p, int SP) {
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * SP];        d[1] = v - p[3 * SP + 1];   d[2] = v - p[2 * SP + 2];   d[3] = v - p[SP + 3];
    d[4] = v - p[3];             d[5] = v - p[-SP + 3];      d[6] = v - p[-2 * SP + 2];  d[7] = v - p[-3 * SP + 1];
    d[8] = v - p[-3 * SP];       d[9] = v - p[-3 * SP - 1];  d[10] = v - p[-2 * SP - 2]; d[11] = v - p[-SP - 3];
    d[12] = v - p[-3];           d[13] = v - p[SP - 3];      d[14] = v - p[2 * SP - 2];  d[15] = v - p[3 * SP - 1];
    int mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn3[k] = __vimin3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        mx3[k] = __vimax3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    }
    int a = -255, b = 255;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        a = max(a, __vimin3_s32(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]));
        b = min(b, __vimax3_s32(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]));
    }
    return max(a, -b);
}

struct FastArgs {
    LevelGeom g;
    const uint8_t* img;      // level plane of image 0
    long long img_stride;    // bytes between images
    int pitch;               // row pitch of the level plane
    int32_t* cell_count;     // [n][cells_total]
    uint32_t* stage;         // [n][stage_total]
    long long cells_total, stage_total;
    int ini_th, min_th;
    int SP;                  // strip pitch (bytes, multiple of 4)
    int strip_bytes, tile_bytes, list_bytes;
};

__global__ void __launch_bounds__(kWarps * 32) fast_cells_kernel(const FastArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    const LevelGeom& g = A.g;
    const int groupsX = (g.nCols + kWarps - 1) / kWarps;
    const int ci = blockIdx.x / groupsX;           // cell row
    const int j0 = (blockIdx.x % groupsX) * kWarps;
    const int iniY = kMinBorder + ci * g.hCell;
    if (iniY >= g.maxBY - 3) return;               // ORBextractor.cpp:794
    const int maxY = min(iniY + g.hCell + 6, g.maxBY);
    const int hROI = maxY - iniY;
    const int ncell = min(kWarps, g.nCols - j0);
    const int xBegin = kMinBorder + j0 * g.wCell;
    const int xEnd = min(kMinBorder + (j0 + ncell) * g.wCell + 6, g.maxBX);
    const int xa = xBegin & ~3, xoff = xBegin - xa;
    const int SP = A.SP;
    uint8_t* strip = smem;
    const uint8_t* img = A.img + (long long)blockIdx.y * A.img_stride;

    // ---- stage the strip: 32-bit loads, rows iniY..maxY, bytes xa..xEnd
    {
        const int nwords = (xEnd - xa + 3) >> 2;
        for (int r = threadIdx.x / 32; r < hROI; r += kWarps) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(img + (long long)(iniY + r) * A.pitch + xa);
            uint32_t* dst = reinterpret_cast<uint32_t*>(strip + r * SP);
            for (int w = threadIdx.x & 31; w < nwords; w += 32) dst[w] = __ldg(src + w);
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= ncell) return;
    const int cj = j0 + warp;
    const int iniX = kMinBorder + cj * g.wCell;
    if (iniX >= g.maxBX - 6) return;               // ORBextractor.cpp:803
    const int maxX = min(iniX + g.wCell + 6, g.maxBX);
    const int wd = (maxX - iniX) - 6, hd = hROI - 6;
    const long long cell = g.cell_base + (long long)ci * g.nCols + cj;
    int32_t* out_count = A.cell_count + (long long)blockIdx.y * A.cells_total + cell;
    uint32_t* out_stage = A.stage + (long long)blockIdx.y * A.stage_total + g.stage_base +
                          ((long long)ci * g.nCols + cj) * g.cell_cap;
    if (wd <= 0 || hd <= 0) return;                // count stays 0 (memset by the launcher)

    uint8_t* tile = smem + A.strip_bytes + warp * (A.tile_bytes + A.list_bytes + 512);
    uint16_t* clist = reinterpret_cast<uint16_t*>(tile + A.tile_bytes);
    uint16_t* queue = reinterpret_cast<uint16_t*>(tile + A.tile_bytes + A.list_bytes);  // ring of 256
    const int TP = g.wCell + 2;
    for (int i = lane; i < (A.tile_bytes >> 2); i += 32) reinterpret_cast<uint32_t*>(tile)[i] = 0;
    __syncwarp();

    const int sx0 = xoff + warp * g.wCell;         // strip column of ROI column 0
    const int t_lo = min(A.ini_th, A.min_th);
    const uint32_t k7 = (uint32_t)(0x7f - min(t_lo, 0x7f)) * 0x01010101u;
    const bool t_big = t_lo >= 0x7f;               // thresholds >= 127: bytes can only pass through bit 7

    int qhead = 0, qcount = 0, ncorner = 0;

    auto process = [&](int nproc) {                // full test on queue[qhead .. qhead+nproc)
        int m = 0, idx = 0;
        if (lane < nproc) {
            idx = queue[(qhead + lane) & 255];
            const int lx = idx & 63, ly = idx >> 6;
            m = arc_measure(strip + (ly + 3) * SP + sx0 + 3 + lx, SP);
            if (m > t_lo) tile[(ly + 1) * TP + lx + 1] = (uint8_t)(m - 1);
        }
        const unsigned b = __ballot_sync(0xffffffffu, lane < nproc && m > t_lo);
        if (lane < nproc && m > t_lo) clist[ncorner + __popc(b & ((1u << lane) - 1))] = (uint16_t)idx;
        ncorner += __popc(b);
        qhead = (qhead + nproc) & 255;
        qcount -= nproc;
    };

    // ---- pass 1 + 2
    const int nw = (wd + 3) >> 2, total = nw * hd;
    for (int base = 0; base < total; base += 32) {
        const int it = base + lane;
        uint32_t surv = 0;
        int row = 0, wx = 0;
        if (it < total) {
            row = it / nw; wx = it - row * nw;
            const int c = (row + 3) * SP + sx0 + 3 + 4 * wx;
            const uint32_t C = ld4(strip, c);
            uint32_t p0 = __vabsdiffu4(C, ld4(strip, c + 3 * SP)), p8 = __vabsdiffu4(C, ld4(strip, c - 3 * SP));
            uint32_t p4 = __vabsdiffu4(C, ld4(strip, c + 3)), p12 = __vabsdiffu4(C, ld4(strip, c - 3));
            if (t_big) { p0 &= 0x80808080u; p8 &= 0x80808080u; p4 &= 0x80808080u; p12 &= 0x80808080u; }
            surv = (gt_bytes(p0, k7) | gt_bytes(p8, k7)) & (gt_bytes(p4, k7) | gt_bytes(p12, k7));
            if (surv) {
                uint32_t p2 = __vabsdiffu4(C, ld4(strip, c + 2 * SP + 2)), p10 = __vabsdiffu4(C, ld4(strip, c - 2 * SP - 2));
                uint32_t p6 = __vabsdiffu4(C, ld4(strip, c - 2 * SP + 2)), p14 = __vabsdiffu4(C, ld4(strip, c + 2 * SP - 2));
                if (t_big) { p2 &= 0x80808080u; p10 &= 0x80808080u; p6 &= 0x80808080u; p14 &= 0x80808080u; }
                surv &= (gt_bytes(p2, k7) | gt_bytes(p10, k7)) & (gt_bytes(p6, k7) | gt_bytes(p14, k7));
            }
            const int valid = min(4, wd - 4 * wx);             // bytes of this word inside the detection area
            if (valid < 4) surv &= (1u << (8 * valid)) - 1u;
        }
        // ordered append of the surviving bytes (lane-major, byte-major == row-major)
        const uint32_t nib = ((surv >> 7) & 1) | ((surv >> 14) & 2) | ((surv >> 21) & 4) | ((surv >> 28) & 8);
        int pre = 0, tot = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const unsigned bal = __ballot_sync(0xffffffffu, (nib >> b) & 1);
            pre += __popc(bal & ((1u << lane) - 1));
            tot += __popc(bal);
        }
        if (nib) {
            int pos = qhead + qcount + pre;
#pragma unroll
            for (int b = 0; b < 4; b++)
                if ((nib >> b) & 1) { queue[pos & 255] = (uint16_t)((row << 6) | (4 * wx + b)); pos++; }
        }
        qcount += tot;
        __syncwarp();
        while (qcount >= 32) { process(32); __syncwarp(); }
    }
    if (qcount > 0) process(qcount);
    __syncwarp();

    // ---- pass 3a: non-max suppression, flag in bit 15, does any keypoint reach iniThFAST?
    bool any_ini = false;
    for (int base = 0; base < ncorner; base += 32) {
        const int i = base + lane;
        bool kp = false; int s = 0;
        if (i < ncorner) {
            const int idx = clist[i];
            const uint8_t* q = tile + ((idx >> 6) + 1) * TP + (idx & 63) + 1;
            s = q[0];
            kp = s > q[-1] && s > q[1] && s > q[-TP - 1] && s > q[-TP] && s > q[-TP + 1] && s > q[TP - 1] &&
                 s > q[TP] && s > q[TP + 1];
            if (kp) clist[i] = (uint16_t)(idx | 0x8000);
        }
        any_ini |= __any_sync(0xffffffffu, kp && s >= A.ini_th);
    }
    __syncwarp();
    // ---- pass 3b: ordered emission
    const int th = any_ini ? A.ini_th : A.min_th;
    int cnt = 0;
    for (int base = 0; base < ncorner; base += 32) {
        const int i = base + lane;
        bool emit = false; int idx = 0, s = 0;
        if (i < ncorner) {
            idx = clist[i];
            if (idx & 0x8000) {
                idx &= 0x7fff;
                s = tile[((idx >> 6) + 1) * TP + (idx & 63) + 1];
                emit = s >= th;
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, emit);
        if (emit) out_stage[cnt + __popc(b & ((1u << lane) - 1))] =
            (uint32_t)((idx & 63) + 3) | ((uint32_t)((idx >> 6) + 3) << 8) | ((uint32_t)s << 16);
        cnt += __popc(b);
    }
    if (lane == 0) *out_count = cnt;
}

}  // namespace

int launch_fast(dsx_ctx* ctx, const uint8_t* images, size_t step, size_t img_stride, int n) {
    StageTimer _t(ctx, 1);
    const ShapePlan& P = ctx->plan;
    DSX_CUDA(cudaMemsetAsync(ctx->ws.cell_count, 0, sizeof(int32_t) * (size_t)n * P.cells_total, ctx->stream));
    for (int l = 0; l < P.nlevels; l++) {
        const LevelGeom& g = P.lv[l];
        if (g.n_cells == 0) continue;
        FastArgs A;
        A.g = g;
        A.img = (l == 0) ? images : ctx->ws.pyr + g.offset;
        A.img_stride = (l == 0) ? (long long)img_stride : P.pyr_bytes;
        A.pitch = (l == 0) ? (int)step : g.pitch;
        A.cell_count = ctx->ws.cell_count; A.stage = ctx->ws.stage;
        A.cells_total = P.cells_total; A.stage_total = P.stage_total;
        A.ini_th = std::min(std::max(ctx->p.ini_th_fast, 0), 255);
        A.min_th = std::min(std::max(ctx->p.min_th_fast, 0), 255);
        A.SP = ((kWarps * g.wCell + 6 + 3 + 3) & ~3) + 8;
        A.strip_bytes = ((A.SP * (g.hCell + 6) + 8) + 15) & ~15;
        A.tile_bytes = (((g.wCell + 2) * (g.hCell + 2)) + 15) & ~15;
        A.list_bytes = ((g.wCell * g.hCell * 2) + 15) & ~15;
        const size_t smem = (size_t)A.strip_bytes + (size_t)kWarps * (A.tile_bytes + A.list_bytes + 512);
        if (smem > 200 * 1024) { set_error("FAST cell too large for shared memory"); return DSX_ERR_INVALID; }
        if (smem > 48 * 1024)
            DSX_CUDA(cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int groupsX = (g.nCols + kWarps - 1) / kWarps;
        dim3 grid(g.nRows * groupsX, n);
        fast_cells_kernel<<<grid, kWarps * 32, smem, ctx->stream>>>(A);
        DSX_LAUNCH_CHECK();
    }
    return DSX_OK;
}

}  // namespace dsx
