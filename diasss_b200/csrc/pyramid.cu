// K1 -- scale pyramid.  Replaces ORBextractor::ComputePyramid (ORBextractor.cpp:1115-1140): level l is
// cv::resize(level l-1, INTER_LINEAR) in OpenCV's 8-bit fixed-point arithmetic (11-bit coefficients):
//   H(y,x)  = S[y][sx]*a0 + S[y][sx+1]*a1                       (int32, a0+a1 = 2048)
//   out     = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// The 19-pixel REFLECT_101 border the reference adds around every level is never read by anything
// downstream (FAST ROIs start at 16, the orientation disc has radius 15 around points >= 19 px inside,
// descriptors use a border-less clone) and is therefore not materialised.
//
// Mapping: one CTA produces a 128 x 32 tile of the destination level.  The source rectangle the tile depends on
// (~156 x 40 bytes at scale 1.2) is staged in shared memory with coalesced 32-bit loads, the horizontal pass runs ONCE
// per (source row, destination column) -- 1.25 evaluations per output pixel instead of the 2 a direct gather makes --
// and keeps (H >> 4) as uint16 in shared memory; the vertical pass combines two of those rows per output pixel and
// writes 4 pixels per 32-bit store (a warp writes 128 contiguous bytes).  Bound: HBM (4.65 B per level-0 pixel over
// the five launches); the integer work per pixel is ~1/3 of the direct form's.
#include "dsx_internal.cuh"

namespace dsx {

namespace {
constexpr int kTW = 128, kTH = 32, kRT = 256;
}

__global__ void __launch_bounds__(kRT)
resize_level_kernel(const uint8_t* __restrict__ src_base, long long src_img_stride, int src_pitch,
                    uint8_t* __restrict__ dst_base, long long dst_img_stride, int dst_pitch, int drows, int dcols,
                    const uint32_t* __restrict__ xtab, const uint32_t* __restrict__ ytab, int SW, int SH) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* S = smem;                                                        // [SH][SW] source rectangle
    uint16_t* H = reinterpret_cast<uint16_t*>(smem + ((SH * SW + 15) & ~15));   // [SH][kTW] horizontal pass, already >> 4
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int x1 = min(x0 + kTW, dcols), y1 = min(y0 + kTH, drows);
    const uint8_t* src = src_base + (long long)blockIdx.z * src_img_stride;
    uint8_t* dst = dst_base + (long long)blockIdx.z * dst_img_stride;
    // both tables are monotone: the tile's source rectangle is spanned by its first and last entries
    const uint32_t xl = __ldg(xtab + x0), xh = __ldg(xtab + x1 - 1), yl = __ldg(ytab + y0), yh = __ldg(ytab + y1 - 1);
    const int a_lo = (xl & 0xffff) & ~3;
    const int nw = ((int)((xh & 0xffff) + (xh >> 31)) - a_lo) / 4 + 1;
    const int r_lo = yl & 0xffff;
    const int nr = (int)((yh & 0xffff) + (yh >> 31)) - r_lo + 1;
    {   // stage: 64 lanes across a row, 4 rows per sweep
        const uint8_t* g = src + (long long)r_lo * src_pitch + a_lo;
        for (int w = tid & 63; w < nw; w += 64)
            for (int r = tid >> 6; r < nr; r += kRT / 64)
                reinterpret_cast<uint32_t*>(S + r * SW)[w] = __ldg(reinterpret_cast<const uint32_t*>(g + (long long)r * src_pitch) + w);
    }
    __syncthreads();
    {   // horizontal: thread = one destination column, walks the staged source rows
        const int c = tid & (kTW - 1);
        const uint32_t xt = __ldg(xtab + min(x0 + c, dcols - 1));
        const int sx = (int)(xt & 0xffff) - a_lo, a1 = (xt >> 16) & 0xfff, a0 = 2048 - a1, inc = xt >> 31;
        const uint8_t* p = S + sx + (tid / kTW) * SW;
        uint16_t* h = H + c + (tid / kTW) * kTW;
#pragma unroll 4
        for (int r = tid / kTW; r < nr; r += kRT / kTW, p += (kRT / kTW) * SW, h += (kRT / kTW) * kTW)
            *h = (uint16_t)((p[0] * a0 + p[inc] * a1) >> 4);
    }
    __syncthreads();
    {   // vertical: thread = 4 adjacent columns, 8 rows apart per sweep; ((b*H) >> 16) is one multiply-high by b << 16
        const int xg = (tid & (kTW / 4 - 1)) * 4;
        if (x0 + xg < dcols)
            for (int y = y0 + tid / (kTW / 4); y < y1; y += kRT / (kTW / 4)) {
                const uint32_t yt = __ldg(ytab + y);
                const int r0 = (int)(yt & 0xffff) - r_lo;
                const uint32_t b1 = (yt & 0x0fff0000u), b0 = (2048u << 16) - b1;      // coefficients << 16
                const uint2 u = *reinterpret_cast<const uint2*>(H + r0 * kTW + xg);
                const uint2 v = *reinterpret_cast<const uint2*>(H + (r0 + (int)(yt >> 31)) * kTW + xg);
                const uint32_t p0 = (__umulhi(b0, u.x & 0xffff) + __umulhi(b1, v.x & 0xffff) + 2) >> 2;
                const uint32_t p1 = (__umulhi(b0, u.x >> 16) + __umulhi(b1, v.x >> 16) + 2) >> 2;
                const uint32_t p2 = (__umulhi(b0, u.y & 0xffff) + __umulhi(b1, v.y & 0xffff) + 2) >> 2;
                const uint32_t p3 = (__umulhi(b0, u.y >> 16) + __umulhi(b1, v.y >> 16) + 2) >> 2;
                // dst_pitch is a multiple of 16 and x of 4: the padded tail of the row absorbs the over-write
                *reinterpret_cast<uint32_t*>(dst + (long long)y * dst_pitch + x0 + xg) = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
            }
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
        const int SW = (((int)std::ceil(kTW * sx) + 2 + 3 + 4) + 3) & ~3;
        const int SH = (int)std::ceil(kTH * sy) + 3;
        const size_t smem = (((size_t)SH * SW + 15) & ~(size_t)15) + (size_t)SH * kTW * 2;
        if (smem > 200 * 1024) { set_error("pyramid: scale factor too large for the staged tile"); return DSX_ERR_INVALID; }
        if (smem > 48 * 1024)
            DSX_CUDA(cudaFuncSetAttribute(resize_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((g.cols + kTW - 1) / kTW, (g.rows + kTH - 1) / kTH, n);
        resize_level_kernel<<<grid, kRT, smem, ctx->stream>>>(src, sstride, spitch, ctx->ws.pyr + g.offset, P.pyr_bytes,
                                                               g.pitch, g.rows, g.cols, P.d_tab + P.xtab_off[l],
                                                               P.d_tab + P.ytab_off[l], SW, SH);
        DSX_LAUNCH_CHECK();
    }
    return DSX_OK;
}

}  // namespace dsx
