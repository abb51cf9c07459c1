// K1 -- scale pyramid.  Replaces ORBextractor::ComputePyramid (ORBextractor.cpp:1115-1140): level l is
// cv::resize(level l-1, INTER_LINEAR) in OpenCV's 8-bit fixed-point arithmetic (11-bit coefficients):
//   H(y,x)  = S[y][sx]*a0 + S[y][sx+1]*a1                       (int32, a0+a1 = 2048)
//   out     = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// The 19-pixel REFLECT_101 border the reference adds around every level is never read by anything
// downstream (FAST ROIs start at 16, the orientation disc has radius 15 around points >= 19 px inside,
// descriptors use a border-less clone) and is therefore not materialised.
//
// Mapping: one CTA of 4 warps produces a 128 x 64 tile of the destination level; warp w owns 16 destination rows, a
// lane owns 4 adjacent destination columns and walks down its rows.  The source rectangle the tile depends on
// (~164 x 80 bytes at scale 1.2) is staged in shared memory with coalesced 32-bit loads, plus a copy shifted by one
// byte, so that the tap pair (S[sx], S[sx+1]) of any column is ONE aligned 16-bit load: the horizontal pass is then a
// single two-way dot product per (source row, destination column),  H = dp2a((a0 | a1 << 16), (S[sx] | S[sx+1] << 8)),
// evaluated once per source row the lane crosses (1.2 evaluations per output pixel; the two rows of (H >> 4) a
// destination row needs stay in registers and the lower one is reused by the next destination row).  The vertical
// pass multiplies on the FMA pipe (b * h < 2^27 fits a 32-bit IMAD), PRMT picks the upper halves -- the separately
// truncated terms (b0*h0) >> 16 and (b1*h1) >> 16 -- of two pixels into one register, one three-input add and one
// shift finish both pixels, and the lane writes its 4 pixels as one 32-bit store (a warp writes 128 contiguous bytes).
// Instruction-issue bound (about 13 per output pixel, half of the first version's); HBM traffic is one read of the
// source level and one write of the destination level.
#include "dsx_internal.cuh"

namespace dsx {

namespace {
constexpr int kTW = 128, kTH = 64, kRT = 128, kRowsPerWarp = kTH / (kRT / 32);
}

template <bool WIDE>      // WIDE: source rows are 16-byte aligned (base and pitch): staged with 128-bit loads and stores
__global__ void __launch_bounds__(kRT, 8)
resize_level_kernel(const uint8_t* __restrict__ src_base, long long src_img_stride, int src_pitch, int scols,
                    uint8_t* __restrict__ dst_base, long long dst_img_stride, int dst_pitch, int drows, int dcols,
                    const uint32_t* __restrict__ xtab, const uint32_t* __restrict__ ytab, int SW, int SH) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* S0 = smem;                                   // [SH][SW] source rectangle (SW multiple of 4)
    uint8_t* S1 = smem + (((size_t)SH * SW + 15) & ~(size_t)15);   // the same bytes shifted down by one: S1[i] = S0[i + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int x1 = min(x0 + kTW, dcols), y1 = min(y0 + kTH, drows);
    const uint8_t* src = src_base + (long long)blockIdx.z * src_img_stride;
    uint8_t* dst = dst_base + (long long)blockIdx.z * dst_img_stride;
    // both tables are monotone: the tile's source rectangle is spanned by its first and last entries
    const uint32_t xl = __ldg(xtab + x0), xh = __ldg(xtab + x1 - 1), yl = __ldg(ytab + y0), yh = __ldg(ytab + y1 - 1);
    const int a_lo = (xl & 0xffff) & (WIDE ? ~15 : ~3);
    // words up to the one holding the right tap of the last column, never past the last word of the source row (the tap
    // S[sx + 1] of a clamped column lies outside the image; its coefficient is 0, any staged byte will do)
    const int nw = min(((int)(xh & 0xffff) + 1 - a_lo) / 4 + 1, (scols + 3) / 4 - a_lo / 4);
    const int r_lo = yl & 0xffff;
    const int nr = (int)((yh & 0xffff) + (yh >> 31)) - r_lo + 1;
    const int W = SW >> 2;
    if (WIDE) {
        // stage, 16 bytes per thread and step: one 128-bit load, the next word for the byte that moves into the shifted
        // copy, two 128-bit stores.  Groups beyond the row's pitch are never requested (ng counts whole groups inside
        // it); the extra word is skipped at the very end of a row's pitch (it would belong to the next row -- or to
        // nobody, after the last row of the last image).
        const uint8_t* g = src + (long long)r_lo * src_pitch + a_lo;
        const int ng = min(nw / 4 + 1, (src_pitch - a_lo) / 16);
        const int total = nr * ng;
        constexpr int kB = 4;
        const int step_r = kRT / ng, step_w = kRT - step_r * ng;
        int r = tid / ng, q = tid - r * ng;
        for (int e = tid; e < total; e += kRT * kB) {
            uint4 v[kB];
            uint32_t nx[kB];
            int so[kB];
#pragma unroll
            for (int i = 0; i < kB; i++) {
                const bool ok = e + i * kRT < total;
                const uint8_t* grow = g + (long long)r * src_pitch;
                v[i] = ok ? __ldg(reinterpret_cast<const uint4*>(grow) + q) : make_uint4(0, 0, 0, 0);
                nx[i] = (ok && a_lo + 16 * (q + 1) < src_pitch) ? __ldg(reinterpret_cast<const uint32_t*>(grow) + 4 * (q + 1)) : 0u;
                so[i] = ok ? r * SW + 16 * q : -1;
                r += step_r; q += step_w;
                if (q >= ng) { q -= ng; r++; }
            }
#pragma unroll
            for (int i = 0; i < kB; i++)
                if (so[i] >= 0) {
                    *reinterpret_cast<uint4*>(S0 + so[i]) = v[i];
                    *reinterpret_cast<uint4*>(S1 + so[i]) = make_uint4(__funnelshift_r(v[i].x, v[i].y, 8), __funnelshift_r(v[i].y, v[i].z, 8),
                                                                        __funnelshift_r(v[i].z, v[i].w, 8), __funnelshift_r(v[i].w, nx[i], 8));
                }
        }
    } else
    {   // stage: the rectangle's words are spread over the CTA in row-major order (consecutive lanes read consecutive
        // words); every word is written twice: as it is (S0) and shifted down by one byte (S1; the byte that moves in
        // comes from the next word of the source row, a second load of the same cache line).  Four words are in flight
        // per thread.  A row's last word may borrow a byte beyond the image, which no tap reads with a non-zero weight.
        const uint8_t* g = src + (long long)r_lo * src_pitch + a_lo;
        const int nw_row = (scols + 3) / 4 - a_lo / 4;          // words of the source row from a_lo on
        const int total = nr * nw;
        constexpr int kB = 4;
        const int step_r = kRT / nw, step_w = kRT - step_r * nw;
        int r = tid / nw, w = tid - r * nw;
        for (int e = tid; e < total; e += kRT * kB) {
            uint32_t v[kB], nx[kB];
            int so[kB];
#pragma unroll
            for (int i = 0; i < kB; i++) {
                const bool ok = e + i * kRT < total;
                const uint32_t* grow = reinterpret_cast<const uint32_t*>(g + (long long)r * src_pitch);
                v[i] = ok ? __ldg(grow + w) : 0u;
                nx[i] = (ok && w + 1 < nw_row) ? __ldg(grow + w + 1) : 0u;
                so[i] = ok ? r * SW + 4 * w : -1;
                r += step_r; w += step_w;
                if (w >= nw) { w -= nw; r++; }
            }
#pragma unroll
            for (int i = 0; i < kB; i++)
                if (so[i] >= 0) {
                    *reinterpret_cast<uint32_t*>(S0 + so[i]) = v[i];
                    *reinterpret_cast<uint32_t*>(S1 + so[i]) = __funnelshift_r(v[i], nx[i], 8);
                }
        }
    }
    __syncthreads();
    const int xg = x0 + 4 * lane;
    const bool live = xg < dcols;            // (lanes beyond the level keep running: the warp shuffles below name every lane)
    // per column: where its tap pair lives (shared-memory byte offset of an aligned 16-bit word) and its coefficients
    int off[4];
    uint32_t coef[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t xt = __ldg(xtab + min(xg + k, dcols - 1));
        const int sx = (int)(xt & 0xffff) - a_lo;
        const uint32_t a1 = (xt >> 16) & 0xfff;
        off[k] = (sx & 1) ? (int)(S1 - S0) + sx - 1 : sx;
        coef[k] = (2048u - a1) | (a1 << 16);
    }
    // The warp walks down its destination rows; lane i holds the table entry of row ya + i.  Row y needs (H >> 4) of the
    // source rows sy and sy + 1 (or sy twice when the level is clamped at the bottom): what the previous row left in
    // registers is reused -- its lower row is this row's upper row five times out of six at scale 1.2 -- and every
    // source row is evaluated once per warp.  All branches are warp-uniform.
    const int ya = y0 + warp * kRowsPerWarp, yb = min(ya + kRowsPerWarp, y1);
    if (ya >= yb) return;
    const uint32_t my_yt = __ldg(ytab + min(ya + (lane & (kRowsPerWarp - 1)), drows - 1));
    uint32_t hu[4], hl[4];          // (H >> 4) of the upper and the lower source row
#pragma unroll
    for (int k = 0; k < 4; k++) hu[k] = hl[k] = 0;
    int row_u = -1, row_l = -1;     // which source rows (relative to r_lo) hu and hl hold
    uint8_t* drow = dst + (long long)ya * dst_pitch + xg;
    auto eval = [&](uint32_t (&h)[4], int row) {
        const uint8_t* srow = S0 + row * SW;
#pragma unroll
        for (int k = 0; k < 4; k++)
            h[k] = __dp2a_lo(coef[k], (uint32_t)*reinterpret_cast<const uint16_t*>(srow + off[k]), 0u) >> 4;
    };
    const int nrow = yb - ya;
    for (int i = 0; i < nrow; i++, drow += dst_pitch) {
        const uint32_t yt = __shfl_sync(0xffffffffu, my_yt, i);
        const int sy = (int)(yt & 0xffff) - r_lo;
        if (sy != row_u) {
            if (sy == row_l) {
#pragma unroll
                for (int k = 0; k < 4; k++) hu[k] = hl[k];
            } else
                eval(hu, sy);
            row_u = sy;
        }
        if (yt >> 31) {             // a lower row exists
            if (sy + 1 != row_l) { eval(hl, sy + 1); row_l = sy + 1; }
        } else if (row_l != sy) {   // clamped at the bottom: both taps are the same source row
#pragma unroll
            for (int k = 0; k < 4; k++) hl[k] = hu[k];
            row_l = sy;
        }
        const uint32_t b1 = (yt >> 16) & 0xfff, b0 = 2048u - b1;
        uint32_t t0[4], t1[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { t0[k] = b0 * hu[k]; t1[k] = b1 * hl[k]; }          // < 2^27: the terms are bits 16..26
        // upper halves of two pixels side by side, both terms added with the rounding constant, then >> 2 per half
        const uint32_t s01 = (__byte_perm(t0[0], t0[1], 0x7632) + __byte_perm(t1[0], t1[1], 0x7632) + 0x00020002u) >> 2;
        const uint32_t s23 = (__byte_perm(t0[2], t0[3], 0x7632) + __byte_perm(t1[2], t1[3], 0x7632) + 0x00020002u) >> 2;
        // dst_pitch is a multiple of 16 and xg of 4: the padded tail of the row absorbs the over-write
        if (live) *reinterpret_cast<uint32_t*>(drow) = __byte_perm(s01, s23, 0x6420);
    }
}

// ---- the TMA-staged form (aligned planes, scale <= ~1.3: every level of the usual 1.2 pyramid) -----------------------------
// One CTA of 4 warps produces a 256 x 32 tile; warp w owns 8 destination rows, a lane 8 adjacent destination columns.
//   * staging: ONE tensor-map box {SW bytes, SH rows} of the source plane, issued by one thread (cp.async.bulk.tensor.3d on
//     the TMA unit; the box starts at a 16-byte aligned column, rows / bytes outside the plane arrive as zeros and are only
//     read with weight 0); all threads then write the copy shifted by one byte, 16 bytes at a time, from shared memory.
//   * horizontal pass: the four tap pairs of 4 adjacent destination columns lie inside ONE 8-byte window of the copy whose
//     parity matches the first column (ShapePlan::xspan4 checks the level's table on the host): two aligned 32-bit loads per
//     group and source row, a PRMT per column picks its pair, IDP.2A multiplies.
//   * the two register sets holding (H >> 4) of the upper and the lower source row swap roles from one destination row to
//     the next (the lower row becomes the upper one five times out of six at scale 1.2): no copies, each source row is
//     evaluated once per warp; 8 pixels leave as one 64-bit store.
// About 0.45 warp instructions per output pixel instead of 1.0 (register-staged form above).
namespace {
#ifndef DSX_PYR_VH
#define DSX_PYR_VH 32
#endif
constexpr int kVW = 256, kVH = DSX_PYR_VH, kVT = 128, kVRows = kVH / (kVT / 32);
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
}

__global__ void __launch_bounds__(kVT, kVH == 32 ? 7 : 4)
resize_level_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ dst_base, long long dst_img_stride,
                        int dst_pitch, int drows, int dcols, const uint32_t* __restrict__ xtab,
                        const uint32_t* __restrict__ ytab, int SW, int SH) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar_word;
    uint8_t* S0 = smem;                                                   // [SH][SW], SW a multiple of 16
    const int s1_off = (SH * SW + 127) & ~127;                            // S1[i] = S0[i + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kVW, y0 = blockIdx.y * kVH;
    const int x1 = min(x0 + kVW, dcols), y1 = min(y0 + kVH, drows);
    uint8_t* dst = dst_base + (long long)blockIdx.z * dst_img_stride;
    const uint32_t xl = __ldg(xtab + x0), yl = __ldg(ytab + y0), yh = __ldg(ytab + y1 - 1);
    const int a_lo = (int)(xl & 0xffff) & ~15;
    const int r_lo = yl & 0xffff;
    const int nr = (int)((yh & 0xffff) + (yh >> 31)) - r_lo + 1;
    const uint32_t bar = smem_addr(&bar_word);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SW * SH)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_addr(S0)), "l"(&tmap), "r"(a_lo >> 2), "r"(r_lo), "r"((int)blockIdx.z), "r"(bar) : "memory");
    }
    // while the box is in flight: where each group of 4 columns finds its window, the byte selectors and the coefficients
    const int xg = x0 + 8 * lane;
    const bool live = xg < dcols;
    int base[2];
    uint32_t sel[8], coef[8];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        int sx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t xt = __ldg(xtab + min(xg + 4 * j + k, dcols - 1));
            sx[k] = (int)(xt & 0xffff) - a_lo;
            const uint32_t a1 = (xt >> 16) & 0xfff;
            coef[4 * j + k] = (2048u - a1) | (a1 << 16);
        }
        const int par = sx[0] & 1;                          // odd first column: the window comes from the shifted copy
        const int A = (sx[0] - par) & ~3;
        base[j] = (par ? s1_off : 0) + A;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = sx[k] - par - A;                   // 0..6 (xspan4)
            sel[4 * j + k] = (uint32_t)(o | ((o + 1) << 4));
        }
    }
    const int ya = y0 + warp * kVRows, yb = min(ya + kVRows, y1);
    const uint32_t my_yt = __ldg(ytab + min(ya + (lane & (kVRows - 1)), drows - 1));   // lane i holds row ya + i
    __syncthreads();                   // (the barrier word is initialised before anybody polls it)
    if (tid < 32) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
    __syncthreads();
    {   // the shifted copy of the rows the tile reads
        const int ng = nr * (SW >> 4);
        for (int g = tid; g < ng; g += kVT) {
            const uint4 v = *reinterpret_cast<const uint4*>(S0 + 16 * g);
            const uint32_t nx = *reinterpret_cast<const uint32_t*>(S0 + 16 * g + 16);
            *reinterpret_cast<uint4*>(S0 + s1_off + 16 * g) = make_uint4(__funnelshift_r(v.x, v.y, 8), __funnelshift_r(v.y, v.z, 8),
                                                                          __funnelshift_r(v.z, v.w, 8), __funnelshift_r(v.w, nx, 8));
        }
    }
    __syncthreads();
    if (ya >= yb) return;
    uint32_t hA[8], hB[8];
#pragma unroll
    for (int k = 0; k < 8; k++) hA[k] = hB[k] = 0;
    int rowA = -1, rowB = -1;          // which source rows (relative to r_lo) the two sets hold
    uint8_t* drow = dst + (long long)ya * dst_pitch + xg;
    auto eval = [&](uint32_t (&h)[8], int row) {
        const uint8_t* srow = S0 + row * SW;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t w0 = *reinterpret_cast<const uint32_t*>(srow + base[j]);
            const uint32_t w1 = *reinterpret_cast<const uint32_t*>(srow + base[j] + 4);
#pragma unroll
            for (int k = 0; k < 4; k++) h[4 * j + k] = __dp2a_lo(coef[4 * j + k], __byte_perm(w0, w1, sel[4 * j + k]), 0u) >> 4;
        }
    };
    auto emit = [&](const uint32_t (&up)[8], const uint32_t (&lo)[8], uint32_t yt) {
        const uint32_t b1 = (yt >> 16) & 0xfff, b0 = 2048u - b1;
        uint32_t out[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            uint32_t t0[4], t1[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { t0[k] = b0 * up[4 * j + k]; t1[k] = b1 * lo[4 * j + k]; }      // < 2^27: the terms are bits 16..26
            const uint32_t s01 = (__byte_perm(t0[0], t0[1], 0x7632) + __byte_perm(t1[0], t1[1], 0x7632) + 0x00020002u) >> 2;
            const uint32_t s23 = (__byte_perm(t0[2], t0[3], 0x7632) + __byte_perm(t1[2], t1[3], 0x7632) + 0x00020002u) >> 2;
            out[j] = __byte_perm(s01, s23, 0x6420);
        }
        // dst_pitch is a multiple of 16 and xg of 8: the padded tail of the row absorbs the over-write
        if (live) *reinterpret_cast<uint2*>(drow) = make_uint2(out[0], out[1]);
        drow += dst_pitch;
    };
    // one destination row: U takes the upper source row, L the lower one (all branches are warp-uniform)
    auto step = [&](uint32_t (&U)[8], int& rowU, uint32_t (&L)[8], int& rowL, int i) {
        const uint32_t yt = __shfl_sync(0xffffffffu, my_yt, i);
        const int sy = (int)(yt & 0xffff) - r_lo;
        if (rowU != sy) { eval(U, sy); rowU = sy; }
        if (yt >> 31) {
            if (rowL != sy + 1) { eval(L, sy + 1); rowL = sy + 1; }
            emit(U, L, yt);
        } else
            emit(U, U, yt);          // clamped at the bottom: both taps are the same source row
    };
    const int nrow = yb - ya;
    for (int i = 0; i < nrow; i += 2) {
        step(hA, rowA, hB, rowB, i);
        if (i + 1 < nrow) step(hB, rowB, hA, rowA, i + 1);
    }
}

int launch_pyramid(dsx_ctx* ctx, const uint8_t* images, size_t step, size_t img_stride, int n) {
    StageTimer _t(ctx, 0);
    const ShapePlan& P = ctx->plan;
    for (int l = 1; l < P.nlevels; l++) {
        const LevelGeom& g = P.lv[l];
        const LevelGeom& gs = P.lv[l - 1];
        const uint8_t* src = (l == 1) ? images : ctx->ws.pyr + gs.offset;
        const long long sstride = (l == 1) ? (long long)img_stride : P.pyr_bytes;
        const int spitch = (l == 1) ? (int)step : gs.pitch;
        // staged source rectangle of a tile: (tile extent) * scale + 2 taps + alignment slack
        const double sx = (double)gs.cols / g.cols, sy = (double)gs.rows / g.rows;
        const int SW = (((int)std::ceil(kTW * sx) + 2 + 15 + 16) + 15) & ~15;      // tap span + alignment slack, whole 16-byte groups
        const int SH = (int)std::ceil(kTH * sy) + 3;
        const size_t smem = 2 * ((((size_t)SH * SW + 15) & ~(size_t)15)) + 16;
        if (smem > 200 * 1024) { set_error("pyramid: scale factor too large for the staged tile"); return DSX_ERR_INVALID; }
        if (smem > 48 * 1024) {
            DSX_CUDA(cudaFuncSetAttribute(resize_level_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DSX_CUDA(cudaFuncSetAttribute(resize_level_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        const bool wide = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch | (uintptr_t)sstride) & 15) == 0;
        uint8_t* dstp = ctx->ws.pyr + g.offset;
        // the TMA-staged form: aligned source planes, 8-byte aligned destination rows, a level whose tap pairs fit the
        // 8-byte windows, and a driver that encodes the tensor map
        bool tma = ctx->pyr_tma && wide && P.xspan4[l] && ((reinterpret_cast<uintptr_t>(dstp) | (uintptr_t)g.pitch | (uintptr_t)P.pyr_bytes) & 7) == 0;
        CUtensorMap tmap;
        const int VSW = (((int)std::ceil(kVW * sx) + 2 + 15 + 16) + 15) & ~15;
        const int VSH = (int)std::ceil(kVH * sy) + 3;
        if (tma) tma = make_plane_map(&tmap, src, spitch, gs.rows, sstride, n, VSW, VSH);
        if (tma) {
            const size_t vsmem = 2 * (size_t)((VSH * VSW + 127) & ~127) + 16;
            if (vsmem > 48 * 1024)
                DSX_CUDA(cudaFuncSetAttribute(resize_level_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
            dim3 vgrid((g.cols + kVW - 1) / kVW, (g.rows + kVH - 1) / kVH, n);
            resize_level_tma_kernel<<<vgrid, kVT, vsmem, ctx->stream>>>(tmap, dstp, P.pyr_bytes, g.pitch, g.rows, g.cols, P.d_tab + P.xtab_off[l],
                                                                         P.d_tab + P.ytab_off[l], VSW, VSH);
            DSX_LAUNCH_CHECK();
            continue;
        }
        dim3 grid((g.cols + kTW - 1) / kTW, (g.rows + kTH - 1) / kTH, n);
        if (wide)
            resize_level_kernel<true><<<grid, kRT, smem, ctx->stream>>>(src, sstride, spitch, gs.cols, dstp, P.pyr_bytes,
                                                                         g.pitch, g.rows, g.cols, P.d_tab + P.xtab_off[l],
                                                                         P.d_tab + P.ytab_off[l], SW, SH);
        else
            resize_level_kernel<false><<<grid, kRT, smem, ctx->stream>>>(src, sstride, spitch, gs.cols, dstp, P.pyr_bytes,
                                                                          g.pitch, g.rows, g.cols, P.d_tab + P.xtab_off[l],
                                                                          P.d_tab + P.ytab_off[l], SW, SH);
        DSX_LAUNCH_CHECK();
    }
    return DSX_OK;
}

}  // namespace dsx
