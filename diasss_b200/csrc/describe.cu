// K4 + K5 + K6 -- orientation, 13x13 Gaussian and rotated rBRIEF for the selected keypoints, then
// operator()'s output assembly and Frame::DetectFeature's mask filter.
//
//   K4  IC_Angle (ORBextractor.cpp:77-104): integer moments over the radius-15 disc (umax rows) of the
//       UNBLURRED level, angle = cv::fastAtan2((float)m01,(float)m10) (OpenCV mathfuncs_core atan_f32; float
//       polynomial, evaluated here with round-to-nearest mul/add, no FMA).
//   K5  cv::GaussianBlur(level, 13x13, sigma 2, REFLECT_101) (:1091-1092) in OpenCV's 8-bit fixed-point form:
//       Q8 kernel [1,2,7,16,31,45,52,45,31,16,7,2,1], H = sum k*P (16 bit), out = (sum k*H + 32768) >> 16.
//       The descriptor only ever reads blurred pixels within 18 px of a keypoint (pattern radius 18.38), so
//       instead of blurring whole levels the kernel blurs one 37x37 window per keypoint from a 49x49 source
//       window staged in shared memory -- identical values, 1/10th of the traffic on large swaths.
//   K6  computeOrbDescriptor (:108-147): a = cosf(angle), b = sinf(angle) as the reference gets them from its C
//       library (glibc 2.39's algorithm, libm_sincosf below: bit-identical to the host libm on every float in
//       [0, 2*pi]), sample = blurred[cvRound(x*b+y*a)][cvRound(x*a-y*b)], 256 comparisons -> 32 bytes.
//   assembly (:1065-1112): level-major order, pt *= mvScaleFactor[level] for level > 0.
//   mask filter (frame.cpp:184-195): keep keypoint iff mask(int(pt.y), int(pt.x)) != 0, order preserved.
//
// One CTA of 128 threads per selected keypoint slot.
#include "dsx_internal.cuh"

namespace dsx {

namespace {

// the 256 test point pairs (ORBextractor.cpp:115-372), 4 signed bytes per test.  In GLOBAL memory: every thread reads its
// own tests, and per-thread addresses serialise on the constant cache (76 % of the kernel's cycles when they lived there);
// as 8- and 16-byte read-only loads they are one L1 request per warp.
__device__ __align__(16) signed char g_pattern[1024] = {
#include "orb_pattern.inc"
};
__constant__ int c_umax[16];
__constant__ int c_gauss[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};

constexpr int kSrcW = 49, kSrcP = 52;   // source window, pitch
constexpr int kBlurW = 37, kHP = 38, kBP = 40;
constexpr int kHalfSrc = 24, kHalfBlur = 18;

struct DescArgs {
    LevelGeom lv[DSX_MAX_LEVELS];
    int nlevels;
    const uint8_t* images; long long img_stride; int step;   // level 0
    const uint8_t* pyr; long long pyr_bytes;                  // levels >= 1
    const uint32_t* key_xy; const uint8_t* key_resp; const int32_t* key_count;
    int keys_total;
    dsx_keypoint* out_kps; uint8_t* out_desc; int32_t* out_count; int cap;
    const uint8_t* blur; long long blur_bytes; long long blur_off[DSX_MAX_LEVELS]; int blur_pitch[DSX_MAX_LEVELS];   // dense form
};

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p;
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// cosf / sinf of glibc 2.39 (the C library the reference links; ORBextractor.cpp:113 reaches them through
// std::cos(float) / std::sin(float)): the Arm Optimized Routines algorithm of sysdeps/ieee754/flt-32/s_sinf.c,
// s_cosf.c, sincosf.h -- argument in double, quadrant by a scaled integer conversion, r = x - n*pi/2, odd / even
// polynomial in double, one rounding to float.  Restated for |x| < 120; this file is compiled with -fmad=false
// and the host oracle's identical restatement is scanned against libm over all 1.09e9 floats of [0, 2*pi].
__device__ __forceinline__ float sincosf_poly(double x, double x2, int tab, int n) {
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = x * x2, t = s2 + x2 * s3, x7 = x3 * x2, s = x + x3 * s1;
        return __double2float_rn(s + x7 * t);
    }
    const double sg = tab ? -1.0 : 1.0;   // the second table is the negated cosine polynomial
    const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5;
    const double c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = x2 * x2, q2 = c3 + x2 * c4, q1 = c0 + x2 * c1, x6 = x4 * x2, cc = q1 + x4 * c2;
    return __double2float_rn(cc + x6 * q2);
}
__device__ __forceinline__ void libm_sincosf(float y, float* sn, float* cs) {
    const unsigned top12 = (__float_as_uint(y) >> 20) & 0x7ffu;
    double x = (double)y;
    if (top12 < 0x3f4u) {                      // |y| < 0.75, compared on the top 12 bits as glibc does
        if (top12 < 0x398u) { *sn = y; *cs = 1.0f; return; }   // |y| < 2^-12
        const double x2 = x * x;
        *sn = sincosf_poly(x, x2, 0, 0);
        *cs = sincosf_poly(x, x2, 0, 1);
        return;
    }
    const double r = x * 0x1.45F306DC9C883p+23;   // 2/pi * 2^24
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = x - (double)n * 0x1.921FB54442D18p0;
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const int tab = (n & 2) ? 1 : 0;
    *sn = sincosf_poly(x * s, x * x, tab, n);
    *cs = sincosf_poly(x * s, x * x, tab, n ^ 1);
}

__global__ void sincosf_probe_kernel(const float* x, float* s, float* c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) libm_sincosf(x[i], s + i, c + i);
}

// Shared-memory layout of the window form.  The 49x49 source window is staged at the 4-byte alignment it has in the
// level plane: smem column so + j holds window column j (so = (kx - 24) & 3), so that rows arrive as aligned 32-bit words
// and the horizontal pass can take four taps per IDP.4A.  Everything downstream is indexed by smem column b = so + c.
constexpr int kSP2 = 64;           // source row pitch (bytes): 13 staged words + slack
constexpr int kHB = 40;            // smem columns the blur passes cover (so + 36 <= 39)
constexpr int kHPairs = 25;        // H rows are stored in vertical PAIRS (row 2p in the low half-word, 2p+1 in the high)

// weights of the Q8 kernel [1,2,7,16,31,45,52,45,31,16,7,2,1] packed for the dot-product instructions
__device__ __forceinline__ uint32_t gauss_w4(int first) {   // bytes g[first .. first+3], zero outside 0..12
    const int g[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int t = first + k; if (t >= 0 && t <= 12) w |= (uint32_t)g[t] << (8 * k); }
    return w;
}
__device__ __forceinline__ uint32_t gauss_w2(int first) {   // bytes g[first], g[first+1]
    const int g[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) { const int t = first + k; if (t >= 0 && t <= 12) w |= (uint32_t)g[t] << (8 * k); }
    return w;
}

__global__ void __launch_bounds__(128) describe_kernel(const DescArgs A) {
    __shared__ __align__(16) uint8_t s_src[kSrcW * kSP2];
    __shared__ __align__(16) uint32_t s_hp[kHPairs * kHB];
    __shared__ __align__(16) uint8_t s_blur[kBlurW * kHB];
    __shared__ int s_m[2][4];
    __shared__ float s_ab[3];
    __shared__ uint8_t s_bits[128];

    const int img = blockIdx.y, slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int level = 0;
    while (level + 1 < A.nlevels && slot >= A.lv[level + 1].key_base) level++;
    const LevelGeom& g = A.lv[level];
    const int i = slot - g.key_base;
    const int32_t* kc = A.key_count + img * DSX_MAX_LEVELS;
    if (slot == 0 && tid == 0) {   // total keypoints of the image, written once
        int tot = 0;
        for (int l = 0; l < A.nlevels; l++) tot += kc[l];
        A.out_count[img] = tot;
    }
    if (i >= kc[level]) return;
    int out_idx = i;
    for (int l = 0; l < level; l++) out_idx += kc[l];
    const uint32_t xy = A.key_xy[(long long)img * A.keys_total + slot];
    const int kx = xy & 0xffff, ky = xy >> 16;
    const uint8_t* plane = (level == 0) ? A.images + (long long)img * A.img_stride : A.pyr + (long long)img * A.pyr_bytes + g.offset;
    const int pitch = (level == 0) ? A.step : g.pitch;
    const int xa = (kx - kHalfSrc) & ~3, so = (kx - kHalfSrc) - xa;      // aligned start of the window's rows; xa may be < 0

    // ---- stage the 49x49 source window.  Inside the level (and with a 4-byte aligned plane): 13 aligned words per
    //      row; otherwise byte by byte with REFLECT_101 at the level border.
    const bool fast_stage = xa >= 0 && kx + kHalfSrc < g.cols && xa + 52 <= pitch && ky - kHalfSrc >= 0 &&
                            ky + kHalfSrc < g.rows && ((reinterpret_cast<uintptr_t>(plane) | (uintptr_t)pitch) & 3) == 0;
    if (fast_stage) {
        const uint8_t* p0 = plane + (long long)(ky - kHalfSrc) * pitch + xa;
        for (int e = tid; e < kSrcW * 13; e += 128) {
            const int r = e / 13, w = e - r * 13;
            reinterpret_cast<uint32_t*>(s_src + r * kSP2)[w] = __ldg(reinterpret_cast<const uint32_t*>(p0 + (long long)r * pitch) + w);
        }
    } else {
        for (int e = tid; e < kSrcW * kSrcW; e += 128) {
            const int r = e / kSrcW, c = e - r * kSrcW;
            const int yy = reflect101(ky - kHalfSrc + r, g.rows), xx = reflect101(kx - kHalfSrc + c, g.cols);
            s_src[r * kSP2 + so + c] = __ldg(plane + (long long)yy * pitch + xx);
        }
    }
    __syncthreads();

    // ---- K4: intensity centroid on the unblurred window (centre at window [24][24])
    {
        int m10 = 0, m01 = 0;
        for (int v = -kHalfPatch + warp; v <= kHalfPatch; v += 4) {
            const int d = c_umax[v < 0 ? -v : v];
            const int u = lane - kHalfPatch;
            if (lane < 31 && u >= -d && u <= d) {
                const int val = s_src[(kHalfSrc + v) * kSP2 + so + kHalfSrc + u];
                m10 += u * val;
                m01 += v * val;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        if (lane == 0) { s_m[0][warp] = m10; s_m[1][warp] = m01; }
    }
    // ---- K5 horizontal pass: H[r][b] = sum_k g[k] * src[r][b + k] for rows 0..48 and smem columns b = 0..39; a thread
    //      takes four adjacent columns from four aligned words, four taps per IDP.4A (the weights of column b + i are the
    //      kernel shifted by i bytes)
    for (int e = tid; e < kSrcW * (kHB / 4); e += 128) {
        const int r = e / (kHB / 4), q = e - r * (kHB / 4);
        const uint32_t* p = reinterpret_cast<const uint32_t*>(s_src + r * kSP2) + q;
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
        uint16_t* hrow = reinterpret_cast<uint16_t*>(s_hp + (r >> 1) * kHB + 4 * q) + (r & 1);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t acc = __dp4a(w0, gauss_w4(0 - i), 0u);
            acc = __dp4a(w1, gauss_w4(4 - i), acc);
            acc = __dp4a(w2, gauss_w4(8 - i), acc);
            acc = __dp4a(w3, gauss_w4(12 - i), acc);
            hrow[2 * i] = (uint16_t)acc;           // <= 255 * 256
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int m10 = s_m[0][0] + s_m[0][1] + s_m[0][2] + s_m[0][3];
        const int m01 = s_m[1][0] + s_m[1][1] + s_m[1][2] + s_m[1][3];
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        const float factorPI = (float)(3.14159265358979323846 / 180.f);      // ORBextractor.cpp:107
        const float rad = __fmul_rn(angle, factorPI);
        libm_sincosf(rad, &s_ab[1], &s_ab[0]);
        s_ab[2] = angle;
    }
    // ---- K5 vertical pass: blurred[rr][b] = (sum_k g[k] * H[rr + k][b] + 32768) >> 16 for rr = 0..36, two taps per
    //      IDP.2A on the vertical pairs; even and odd rows in separate sweeps so that the weights are constants
    for (int e = tid; e < 19 * kHB; e += 128) {               // rr = 0, 2, .., 36: pairs rr/2 .. rr/2 + 6
        const int h = e / kHB, b = e - h * kHB;
        const uint32_t* p = s_hp + h * kHB + b;
        uint32_t acc = 32768u;
#pragma unroll
        for (int k = 0; k < 7; k++) acc = __dp2a_lo(p[k * kHB], gauss_w2(2 * k), acc);
        s_blur[(2 * h) * kHB + b] = (uint8_t)(acc >> 16);
    }
    for (int e = tid; e < 18 * kHB; e += 128) {               // rr = 1, 3, .., 35: pairs (rr - 1)/2 .. + 6
        const int h = e / kHB, b = e - h * kHB;
        const uint32_t* p = s_hp + h * kHB + b;
        uint32_t acc = 32768u;
#pragma unroll
        for (int k = 0; k < 7; k++) acc = __dp2a_lo(p[k * kHB], gauss_w2(2 * k - 1), acc);
        s_blur[(2 * h + 1) * kHB + b] = (uint8_t)(acc >> 16);
    }
    __syncthreads();

    // ---- K6: two tests per thread
    {
        const float a = s_ab[0], b = s_ab[1];
        int bits = 0;
#pragma unroll
        const uint2 pw = __ldg(reinterpret_cast<const uint2*>(g_pattern) + tid);      // tests 2 tid and 2 tid + 1
        const uint32_t pword[2] = {pw.x, pw.y};
        for (int e = 0; e < 2; e++) {
            const signed char pt[4] = {(signed char)(pword[e] & 0xff), (signed char)((pword[e] >> 8) & 0xff),
                                       (signed char)((pword[e] >> 16) & 0xff), (signed char)(pword[e] >> 24)};
            int v[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float px = (float)pt[2 * q], py = (float)pt[2 * q + 1];
                const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
                const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
                v[q] = s_blur[(kHalfBlur + iy) * kHB + so + kHalfBlur + ix];
            }
            bits |= (v[0] < v[1]) << e;
        }
        s_bits[tid] = (uint8_t)bits;
    }
    __syncthreads();
    const long long o = (long long)img * A.cap + out_idx;
    if (tid < 32) {
        const int byte = s_bits[4 * tid] | (s_bits[4 * tid + 1] << 2) | (s_bits[4 * tid + 2] << 4) | (s_bits[4 * tid + 3] << 6);
        A.out_desc[o * 32 + tid] = (uint8_t)byte;
    }
    if (tid == 32) {
        dsx_keypoint kp;
        kp.x = (level != 0) ? __fmul_rn((float)kx, g.scale) : (float)kx;      // :1103-1109
        kp.y = (level != 0) ? __fmul_rn((float)ky, g.scale) : (float)ky;
        kp.size = g.kp_size;
        kp.angle = s_ab[2];
        kp.response = (float)A.key_resp[(long long)img * A.keys_total + slot];
        kp.octave = level;
        kp.class_id = -1;
        A.out_kps[o] = kp;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Dense form of K5 / K6 (many keypoints per image, BASELINE config 5): every level is blurred ONCE
// (cv::GaussianBlur 13x13 sigma 2 REFLECT_101 in the same fixed-point arithmetic), then a warp per keypoint takes its
// moments from the unblurred level and its 512 samples from the blurred plane.  Same bytes as the per-keypoint form.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBT = 32, kBW = 128;     // blur tile: 32 rows x 128 columns per CTA of 256 threads

struct BlurArgs {
    LevelGeom g;
    const uint8_t* src; long long src_stride; int src_pitch;
    uint8_t* dst; long long dst_stride; int dst_pitch;
};

__global__ void __launch_bounds__(256) blur_level_kernel(const BlurArgs A) {
    __shared__ __align__(16) uint8_t s_in[(kBT + 12) * (kBW + 16)];
    __shared__ __align__(16) uint16_t s_hh[(kBT + 12) * kBW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kBW, y0 = blockIdx.y * kBT;
    const int rows = A.g.rows, cols = A.g.cols;
    const uint8_t* src = A.src + (long long)blockIdx.z * A.src_stride;
    uint8_t* dst = A.dst + (long long)blockIdx.z * A.dst_stride;
    constexpr int SWI = kBW + 16;
    // stage rows y0-6 .. y0+kBT+5, columns x0-6 .. x0+kBW+5, reflected at the level border
    for (int e = tid; e < (kBT + 12) * (kBW + 12); e += 256) {
        const int r = e / (kBW + 12), c = e - r * (kBW + 12);
        const int yy = reflect101(min(y0 - 6 + r, rows + 5), rows), xx = reflect101(min(x0 - 6 + c, cols + 5), cols);
        s_in[r * SWI + c] = __ldg(src + (long long)yy * A.src_pitch + xx);
    }
    __syncthreads();
    for (int e = tid; e < (kBT + 12) * kBW; e += 256) {
        const int r = e / kBW, c = e - r * kBW;
        const uint8_t* p = s_in + r * SWI + c;
        int acc = 0;
#pragma unroll
        for (int k = 0; k < 13; k++) acc += c_gauss[k] * p[k];
        s_hh[r * kBW + c] = (uint16_t)acc;
    }
    __syncthreads();
    for (int e = tid; e < kBT * kBW; e += 256) {
        const int r = e / kBW, c = e - r * kBW;
        if (y0 + r >= rows || x0 + c >= cols) continue;
        unsigned acc = 32768u;
#pragma unroll
        for (int k = 0; k < 13; k++) acc += (unsigned)c_gauss[k] * s_hh[(r + k) * kBW + c];
        dst[(long long)(y0 + r) * A.dst_pitch + x0 + c] = (uint8_t)(acc >> 16);
    }
}

// one warp per selected keypoint slot, 4 per CTA
__global__ void __launch_bounds__(128) describe_dense_kernel(const DescArgs A) {
    const int img = blockIdx.y, slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (slot >= A.keys_total) return;
    int level = 0;
    while (level + 1 < A.nlevels && slot >= A.lv[level + 1].key_base) level++;
    const LevelGeom& g = A.lv[level];
    const int i = slot - g.key_base;
    const int32_t* kc = A.key_count + img * DSX_MAX_LEVELS;
    if (slot == 0 && lane == 0) {   // total keypoints of the image, written once
        int tot = 0;
        for (int l = 0; l < A.nlevels; l++) tot += kc[l];
        A.out_count[img] = tot;
    }
    if (i >= kc[level]) return;
    int out_idx = i;
    for (int l = 0; l < level; l++) out_idx += kc[l];
    const uint32_t xy = A.key_xy[(long long)img * A.keys_total + slot];
    const int kx = xy & 0xffff, ky = xy >> 16;
    const uint8_t* plane = (level == 0) ? A.images + (long long)img * A.img_stride : A.pyr + (long long)img * A.pyr_bytes + g.offset;
    const int pitch = (level == 0) ? A.step : g.pitch;
    // K4: intensity centroid, lane = u + 15 (keypoints lie >= 19 px inside the level: no border handling)
    int m10 = 0, m01 = 0;
    {
        const int u = lane - kHalfPatch;
        const uint8_t* p = plane + (long long)ky * pitch + kx + u;
#pragma unroll 4
        for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
            const int d = c_umax[v < 0 ? -v : v];
            if (lane < 31 && u >= -d && u <= d) {
                const int val = __ldg(p + (long long)v * pitch);
                m10 += u * val;
                m01 += v * val;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    const float factorPI = (float)(3.14159265358979323846 / 180.f);      // ORBextractor.cpp:107
    float a, b;
    libm_sincosf(__fmul_rn(angle, factorPI), &b, &a);
    // K6: lane = descriptor byte, 8 tests each, samples from the blurred plane
    const uint8_t* bl = A.blur + (long long)img * A.blur_bytes + A.blur_off[level] + (long long)ky * A.blur_pitch[level] + kx;
    const int bp = A.blur_pitch[level];
    int bits = 0;
    const uint4 pw0 = __ldg(reinterpret_cast<const uint4*>(g_pattern) + 2 * lane), pw1 = __ldg(reinterpret_cast<const uint4*>(g_pattern) + 2 * lane + 1);
    const uint32_t pword[8] = {pw0.x, pw0.y, pw0.z, pw0.w, pw1.x, pw1.y, pw1.z, pw1.w};      // tests 8 lane .. 8 lane + 7
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const signed char pt[4] = {(signed char)(pword[e] & 0xff), (signed char)((pword[e] >> 8) & 0xff),
                                   (signed char)((pword[e] >> 16) & 0xff), (signed char)(pword[e] >> 24)};
        int v[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const float px = (float)pt[2 * q], py = (float)pt[2 * q + 1];
            const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            v[q] = __ldg(bl + (long long)iy * bp + ix);
        }
        bits |= (v[0] < v[1]) << e;
    }
    const long long o = (long long)img * A.cap + out_idx;
    A.out_desc[o * 32 + lane] = (uint8_t)bits;
    if (lane == 0) {
        dsx_keypoint kp;
        kp.x = (level != 0) ? __fmul_rn((float)kx, g.scale) : (float)kx;      // :1103-1109
        kp.y = (level != 0) ? __fmul_rn((float)ky, g.scale) : (float)ky;
        kp.size = g.kp_size;
        kp.angle = angle;
        kp.response = (float)A.key_resp[(long long)img * A.keys_total + slot];
        kp.octave = level;
        kp.class_id = -1;
        A.out_kps[o] = kp;
    }
}

// Frame::DetectFeature's filter: ordered compaction of one image's keypoints by the mask.
__global__ void __launch_bounds__(1024) finalize_kernel(const dsx_keypoint* __restrict__ in_kps, const uint8_t* __restrict__ in_desc,
                                                        const int32_t* __restrict__ in_count, int in_cap,
                                                        const uint8_t* __restrict__ masks, long long mstep, long long mask_stride,
                                                        dsx_keypoint* out_kps, uint8_t* out_desc, int32_t* out_count, int out_cap,
                                                        int32_t* err_flag) {
    __shared__ int wsum[33];
    __shared__ int s_base;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = in_count[img];
    const dsx_keypoint* kps = in_kps + (long long)img * in_cap;
    const uint4* desc = reinterpret_cast<const uint4*>(in_desc + (long long)img * in_cap * 32);
    const uint8_t* mask = masks ? masks + (long long)img * mask_stride : nullptr;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        bool keep = false;
        dsx_keypoint kp;
        if (i < n) {
            kp = kps[i];
            keep = mask ? (mask[(long long)(int)kp.y * mstep + (int)kp.x] != 0) : true;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int wcount = __popc(bal);
        if (lane == 0) wsum[warp] = wcount;
        __syncthreads();
        if (warp == 0) {
            int v = wsum[lane], incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            wsum[lane] = incl - v;
            if (lane == 31) wsum[32] = incl;
        }
        __syncthreads();
        const int pos = s_base + wsum[warp] + __popc(bal & ((1u << lane) - 1));
        if (keep && pos < out_cap) {
            out_kps[(long long)img * out_cap + pos] = kp;
            uint4* od = reinterpret_cast<uint4*>(out_desc + ((long long)img * out_cap + pos) * 32);
            od[0] = desc[2 * i];
            od[1] = desc[2 * i + 1];
        }
        __syncthreads();
        if (tid == 0) s_base += wsum[32];
        __syncthreads();
    }
    if (tid == 0) {
        out_count[img] = min(s_base, out_cap);
        if (s_base > out_cap) atomicExch(err_flag, DSX_ERR_CAPACITY);
    }
}

// Per-keypoint Frame::geo_img look-up (FEAmatcher.cpp:81-82) evaluated from the per-ping model
// (frame.cpp:134-151): geo = pose + g_range[k] * cos/sin(yaw +- PI/2), round-to-nearest mul then add.
__global__ void georef_kernel(const dsx_keypoint* __restrict__ kps, const int32_t* __restrict__ count, int cap,
                              const double* __restrict__ rowtab6, const double* __restrict__ g_range, int rows, int cols,
                              int n_range, double* __restrict__ geo_xy) {
    const int img = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count[img]) return;
    const dsx_keypoint kp = kps[(long long)img * cap + i];
    const int r = (int)kp.y, c = (int)kp.x;
    const double* t = rowtab6 + ((long long)img * rows + r) * 6;
    const double* g = g_range + (long long)img * n_range;
    const int half = cols / 2;
    double gx, gy;
    if (c >= half) {
        const double gr = g[c - half];
        gx = __dadd_rn(t[0], __dmul_rn(gr, t[2]));
        gy = __dadd_rn(t[1], __dmul_rn(gr, t[3]));
    } else {
        const double gr = g[(cols - half) - c];
        gx = __dadd_rn(t[0], __dmul_rn(gr, t[4]));
        gy = __dadd_rn(t[1], __dmul_rn(gr, t[5]));
    }
    geo_xy[((long long)img * cap + i) * 2] = gx;
    geo_xy[((long long)img * cap + i) * 2 + 1] = gy;
}

}  // namespace

int upload_umax(dsx_ctx* ctx) {
    DSX_CUDA(cudaMemcpyToSymbolAsync(c_umax, ctx->umax, sizeof(int) * 16, 0, cudaMemcpyHostToDevice, ctx->stream));
    return DSX_OK;
}

int launch_describe(dsx_ctx* ctx, const uint8_t* images, size_t step, size_t img_stride, int n) {
    StageTimer _t(ctx, 3);
    const ShapePlan& P = ctx->plan;
    DescArgs A;
    for (int l = 0; l < P.nlevels; l++) A.lv[l] = P.lv[l];
    A.nlevels = P.nlevels;
    A.images = images; A.img_stride = (long long)img_stride; A.step = (int)step;
    A.pyr = ctx->ws.pyr; A.pyr_bytes = P.pyr_bytes;
    A.key_xy = ctx->ws.key_xy; A.key_resp = ctx->ws.key_resp; A.key_count = ctx->ws.key_count;
    A.keys_total = P.keys_total;
    A.out_kps = ctx->ws.tmp_kps; A.out_desc = ctx->ws.tmp_desc; A.out_count = ctx->ws.tmp_count; A.cap = ctx->cap;
    A.blur = nullptr; A.blur_bytes = 0;
    if (P.dense_describe) {
        A.blur = ctx->ws.blur; A.blur_bytes = P.blur_bytes;
        for (int l = 0; l < P.nlevels; l++) {
            A.blur_off[l] = P.blur_off[l]; A.blur_pitch[l] = P.blur_pitch[l];
            BlurArgs B;
            B.g = P.lv[l];
            B.src = (l == 0) ? images : ctx->ws.pyr + P.lv[l].offset;
            B.src_stride = (l == 0) ? (long long)img_stride : P.pyr_bytes;
            B.src_pitch = (l == 0) ? (int)step : P.lv[l].pitch;
            B.dst = ctx->ws.blur + P.blur_off[l]; B.dst_stride = P.blur_bytes; B.dst_pitch = P.blur_pitch[l];
            dim3 bg((P.lv[l].cols + kBW - 1) / kBW, (P.lv[l].rows + kBT - 1) / kBT, n);
            blur_level_kernel<<<bg, 256, 0, ctx->stream>>>(B);
            DSX_LAUNCH_CHECK();
        }
        dim3 grid((P.keys_total + 3) / 4, n);
        describe_dense_kernel<<<grid, 128, 0, ctx->stream>>>(A);
        DSX_LAUNCH_CHECK();
        return DSX_OK;
    }
    dim3 grid(P.keys_total, n);
    describe_kernel<<<grid, 128, 0, ctx->stream>>>(A);
    DSX_LAUNCH_CHECK();
    return DSX_OK;
}

int launch_sincosf_probe(dsx_ctx* ctx, const float* x, float* s, float* c, int n) {
    sincosf_probe_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(x, s, c, n);
    DSX_LAUNCH_CHECK();
    return DSX_OK;
}

int launch_finalize(dsx_ctx* ctx, const uint8_t* masks, size_t mstep, size_t mask_stride, int n, int rows, int cols,
                    dsx_keypoint* out_kps, uint8_t* out_desc, int32_t* out_count, int out_cap) {
    StageTimer _t(ctx, 4);
    (void)rows; (void)cols;
    finalize_kernel<<<n, 1024, 0, ctx->stream>>>(ctx->ws.tmp_kps, ctx->ws.tmp_desc, ctx->ws.tmp_count, ctx->cap, masks,
                                                 (long long)mstep, (long long)mask_stride, out_kps, out_desc, out_count,
                                                 out_cap, ctx->ws.err_flag);
    DSX_LAUNCH_CHECK();
    return DSX_OK;
}

int launch_georef(dsx_ctx* ctx, const dsx_features_dev* f, const double* rowtab6, const double* g_range, int rows,
                  int cols, int n_range) {
    StageTimer _t(ctx, 5);
    dim3 grid((f->cap + 255) / 256, f->n_images);
    georef_kernel<<<grid, 256, 0, ctx->stream>>>(f->kps, f->count, f->cap, rowtab6, g_range, rows, cols, n_range, f->geo_xy);
    DSX_LAUNCH_CHECK();
    return DSX_OK;
}

}  // namespace dsx
