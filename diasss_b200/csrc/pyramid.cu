// K1 -- scale pyramid.  Replaces ORBextractor::ComputePyramid (ORBextractor.cpp:1115-1140): level l is
// cv::resize(level l-1, INTER_LINEAR) in OpenCV's 8-bit fixed-point arithmetic (11-bit coefficients):
//   H(y,x)  = S[y][sx]*a0 + S[y][sx+1]*a1                       (int32, a0+a1 = 2048)
//   out     = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// The 19-pixel REFLECT_101 border the reference adds around every level is never read by anything
// downstream (FAST ROIs start at 16, the orientation disc has radius 15 around points >= 19 px inside,
// descriptors use a border-less clone) and is therefore not materialised.
//
// Mapping: one thread produces 4 horizontally adjacent output pixels (one 32-bit store); a warp writes
// 128 contiguous bytes.  Source bytes are read through the read-only path; the two source rows of a warp
// span ~154 contiguous bytes each, so every fetched sector is fully used.  Bound: HBM (4.65 B per level-0
// pixel over the five launches).
#include "dsx_internal.cuh"

namespace dsx {

__global__ void __launch_bounds__(256)
resize_level_kernel(const uint8_t* __restrict__ src_base, long long src_img_stride, int src_pitch,
                    uint8_t* __restrict__ dst_base, long long dst_img_stride, int dst_pitch, int drows, int dcols,
                    const uint32_t* __restrict__ xtab, const uint32_t* __restrict__ ytab) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= dcols) return;
    const uint8_t* src = src_base + (long long)blockIdx.z * src_img_stride;
    uint8_t* dst = dst_base + (long long)blockIdx.z * dst_img_stride;
    const uint32_t yt = __ldg(ytab + y);
    const int sy0 = yt & 0xffff, b1 = (yt >> 16) & 0xfff, b0 = 2048 - b1;
    const uint8_t* S0 = src + (long long)sy0 * src_pitch;
    const uint8_t* S1 = S0 + (long long)(yt >> 31) * src_pitch;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = min(x4 + k, dcols - 1);
        const uint32_t xt = __ldg(xtab + x);
        const int sx = xt & 0xffff, a1 = (xt >> 16) & 0xfff, a0 = 2048 - a1, inc = xt >> 31;
        const int h0 = __ldg(S0 + sx) * a0 + __ldg(S0 + sx + inc) * a1;
        const int h1 = __ldg(S1 + sx) * a0 + __ldg(S1 + sx + inc) * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out |= (uint32_t)v << (8 * k);
    }
    // dst_pitch is a multiple of 16 and x4 of 4: the padded tail of the row absorbs the over-write
    *reinterpret_cast<uint32_t*>(dst + (long long)y * dst_pitch + x4) = out;
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
        dim3 grid((g.cols + 1023) / 1024, g.rows, n);
        resize_level_kernel<<<grid, 256, 0, ctx->stream>>>(src, sstride, spitch, ctx->ws.pyr + g.offset, P.pyr_bytes,
                                                            g.pitch, g.rows, g.cols, P.d_tab + P.xtab_off[l],
                                                            P.d_tab + P.ytab_off[l]);
        DSX_LAUNCH_CHECK();
    }
    return DSX_OK;
}

}  // namespace dsx
