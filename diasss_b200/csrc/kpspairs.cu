// SURVEY.md section 8f rank 3, first half -- the step AFTER the path: Optimizer::GetKpsPairs
// (src/core/optimizer.cpp:575-639, USE_ANNO = 0 branch), which turns the rows RobustMatching appended to
// Frame::corres_kps into what the GTSAM stage consumes: per kept correspondence the 7-vector
//   [y_s, x_s, slant range_s, y_t, x_t, slant range_t, draping depth = 0]
// with integer-truncated pixel coordinates, correspondences within 20 bins of the nadir line dropped, and
//   slant range = sqrt(altitude[y]^2 + ground_range[|x - n_range|]^2)          (:617-620, double, no FMA).
// One CTA per image pair, ordered compaction (the reference pushes in row order).  The per-correspondence LM solves
// that follow (LoopClosingTFs, :641-982) are GTSAM numerics and stay on the host.
#include "dsx_internal.cuh"

namespace dsx {

namespace {

__global__ void __launch_bounds__(256) kps_pairs_kernel(const double* __restrict__ rows6, const int32_t* __restrict__ cnt,
                                                        const int32_t* __restrict__ off, const int32_t* __restrict__ pairs,
                                                        const int32_t* __restrict__ img_id, const double* __restrict__ alt,
                                                        long long alt_stride, const double* __restrict__ gra, long long gra_stride,
                                                        int n_range, double* __restrict__ out7, int32_t* __restrict__ out_cnt) {
    __shared__ int wsum[8];
    __shared__ int s_base;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = cnt[p];
    const long long o = off[p];
    const int s = pairs[2 * p], t = pairs[2 * p + 1];
    const int id_t = img_id[t];
    const double* alt_s = alt + (long long)s * alt_stride; const double* alt_t = alt + (long long)t * alt_stride;
    const double* gra_s = gra + (long long)s * gra_stride; const double* gra_t = gra + (long long)t * gra_stride;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < K; b0 += 256) {
        const int i = b0 + tid;
        bool keep = false;
        int ys = 0, xs = 0, yt = 0, xt = 0, gs = 0, gt = 0;
        if (i < K) {
            const double* r = rows6 + (o + i) * 6;
            const int id_check = (int)r[1];                                    // :596
            ys = (int)r[2]; xs = (int)r[3]; yt = (int)r[4]; xt = (int)r[5];   // :597-598
            gs = xs - n_range; gt = xt - n_range;                              // :603-604
            keep = !(abs(gs) < 20 || abs(gt) < 20) && id_check == id_t;        // :605-613
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; w++) before += wsum[w];
        if (keep) {
            const int e = before + __popc(bal & ((1u << lane) - 1));
            const double as = alt_s[ys], g1 = gra_s[abs(gs)], at = alt_t[yt], g2 = gra_t[abs(gt)];
            double* q = out7 + (o + e) * 7;
            q[0] = (double)ys; q[1] = (double)xs;
            q[2] = __dsqrt_rn(__dadd_rn(__dmul_rn(as, as), __dmul_rn(g1, g1)));  // :617
            q[3] = (double)yt; q[4] = (double)xt;
            q[5] = __dsqrt_rn(__dadd_rn(__dmul_rn(at, at), __dmul_rn(g2, g2)));  // :619
            q[6] = 0.0;                                                          // drap_depth (USE_ANNO = 0)
        }
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < 8; w++) tot += wsum[w]; s_base += tot; }
        __syncthreads();
    }
    if (tid == 0) out_cnt[p] = s_base;
}

}  // namespace

int launch_kps_pairs(dsx_ctx* ctx, const double* rows6, const int32_t* cnt, const int32_t* off, const int32_t* d_pairs, const int32_t* d_img_id,
                     int n_pairs, const double* alt, long long alt_stride, const double* gra, long long gra_stride, int n_range,
                     double* out7, int32_t* out_cnt) {
    if (n_pairs <= 0) return DSX_OK;
    kps_pairs_kernel<<<n_pairs, 256, 0, ctx->stream>>>(rows6, cnt, off, d_pairs, d_img_id, alt, alt_stride, gra, gra_stride, n_range, out7, out_cnt);
    DSX_LAUNCH_CHECK();
    return DSX_OK;
}

}  // namespace dsx
