// TEST INFRASTRUCTURE ONLY -- CPU oracle for the diasss ORB front end ("ORB mode").
//
// This file is a from-scratch CPU restatement of the reference extractor
//   /root/reference/thirdparty/ORBextractor.cpp  (ORB-SLAM2 fork used by halajun/diasss)
// and of the OpenCV primitives it calls (OpenCV itself is a third-party dependency that is
// not vendored in the reference; "Tested with OpenCV 4.6", README.md:28-29).  Nothing in
// the product (diasss_b200/, include/) may include, link or call this file: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as
// the checker.
//
// Pinning status: PINNED.  The OpenCV primitives (resize, FAST, GaussianBlur, fastAtan2, RNG) are pinned
// bit-exactly against the real OpenCV 4.13 (cv2) in tests/test_oracle_primitives.py and through
// tests/golden/*.npz.  The diasss/ORB-SLAM2 control flow above those primitives (cell grid, quadtree
// distribution, descriptors, assembly) is pinned byte for byte against oracle/_ref -- the reference's OWN
// ORBextractor.cpp compiled unmodified (plus the ORB-mode switch S1) by oracle/build_ref.sh against an
// OpenCV stand-in whose primitives are these very functions -- in tests/test_ref_pin.py: candidates per
// level, pyramid levels, keypoints including order, descriptors, DistributeOctTree on hand-made lists.
// A second, independent restatement on the real OpenCV (oracle/cv2_oracle.py) cross-checks both.
//
// Documented deviations where the reference is undefined:
//   B1  nIni = max(1, round(w/h))              (ORBextractor.cpp:543 divides by nIni==0)
//   B2  final-phase sort tie-break = creation sequence instead of heap address (:684): what the reference build
//       itself produces when its list nodes get addresses in creation order (fresh heap); with glibc's heap the
//       reference's output varies from call to call (tools/ref_tiebreak_stats.py)
//   S1  descriptors: the dormant 4-argument computeDescriptors (rBRIEF) (:1097)
//   A5  cosf/sinf of the keypoint angle (:113) = glibc 2.39's algorithm, restated below (libm_sincosf); equal to the
//       platform libm the reference links for EVERY float in [0, 2*pi] (tests/test_ref_pin.py scans all 1.09e9)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "oracle_capi.h"

namespace {

const int PATCH_SIZE = 31;       // ORBextractor.cpp:72
const int HALF_PATCH_SIZE = 15;  // :73
const int EDGE_THRESHOLD = 19;   // :74

const signed char kPattern[1024] = {
#include "orb_pattern.inc"
};

inline int cvRoundF(float v) { return (int)lrintf(v); }   // cvRound: cvtss2si, half to even
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (v < (float)i); }

struct Image {
    int rows = 0, cols = 0;
    std::vector<uint8_t> d;
    Image() {}
    Image(int r, int c) : rows(r), cols(c), d((size_t)r * c) {}
    uint8_t* row(int y) { return d.data() + (size_t)y * cols; }
    const uint8_t* row(int y) const { return d.data() + (size_t)y * cols; }
};

// ---------------------------------------------------------------------------------------
// cv::resize, 8UC1, INTER_LINEAR (OpenCV imgproc/resize.cpp: HResizeLinear + VResizeLinear
// with INTER_RESIZE_COEF_BITS = 11).  Call site: ORBextractor.cpp:1128.
// ---------------------------------------------------------------------------------------
void resize_linear_u8(const uint8_t* src, int srows, int scols, size_t sstep, uint8_t* dst,
                      int drows, int dcols, size_t dstep) {
    const double inv_scale_x = (double)dcols / scols, inv_scale_y = (double)drows / srows;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    std::vector<int> xofs(dcols), yofs(drows);
    std::vector<short> alpha(2 * (size_t)dcols), beta(2 * (size_t)drows);
    int xmax = dcols;
    for (int dx = 0; dx < dcols; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloorF(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx + 1 >= scols) {
            xmax = std::min(xmax, dx);
            if (sx >= scols - 1) { fx = 0; sx = scols - 1; }
        }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cvRoundF((1.f - fx) * 2048.f);
        alpha[2 * dx + 1] = (short)cvRoundF(fx * 2048.f);
    }
    for (int dy = 0; dy < drows; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloorF(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = (short)cvRoundF((1.f - fy) * 2048.f);
        beta[2 * dy + 1] = (short)cvRoundF(fy * 2048.f);
    }
    std::vector<int> h0(dcols), h1(dcols);
    auto hrow = [&](int sy, std::vector<int>& H) {
        const uint8_t* S = src + (size_t)sy * sstep;
        int dx = 0;
        for (; dx < xmax; dx++) {
            int sx = xofs[dx];
            H[dx] = S[sx] * alpha[2 * dx] + S[sx + 1] * alpha[2 * dx + 1];
        }
        for (; dx < dcols; dx++) H[dx] = S[xofs[dx]] * 2048;
    };
    for (int dy = 0; dy < drows; dy++) {
        int sy0 = std::min(std::max(yofs[dy], 0), srows - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), srows - 1);
        hrow(sy0, h0);
        hrow(sy1, h1);
        const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int x = 0; x < dcols; x++)
            D[x] = (uint8_t)((((b0 * (h0[x] >> 4)) >> 16) + ((b1 * (h1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------------------------------
// cv::FAST(roi, kps, threshold, nonmaxSuppression=true)  == FAST_t<16> (TYPE_9_16)
// (OpenCV features2d/fast.cpp, fast_score.cpp).  Call sites: ORBextractor.cpp:809,814.
// ---------------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// m(p): largest over the 16 arcs of 9 consecutive ring pixels of min(d) / min(-d).
inline int fast_arc_measure(const uint8_t* p, size_t step) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[(ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -255;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}

struct RawKp { int x, y, score; };

void fast9_16(const uint8_t* img, int rows, int cols, size_t step, int threshold,
              std::vector<RawKp>& out) {
    out.clear();
    if (rows < 7 || cols < 7) return;
    std::vector<uint8_t> score((size_t)rows * cols, 0);
    const ptrdiff_t st = (ptrdiff_t)step;
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            const uint8_t* p = img + (size_t)y * step + x;
            // OpenCV-style early rejection (fast.cpp): a 9-arc holds one pixel of every opposite ring pair
            const int hi = p[0] + threshold, lo = p[0] - threshold;
            const int a0 = p[3 * st], a8 = p[-3 * st], a4 = p[3], a12 = p[-3];
            bool br = (a0 > hi || a8 > hi) && (a4 > hi || a12 > hi);
            bool dk = (a0 < lo || a8 < lo) && (a4 < lo || a12 < lo);
            if (!(br || dk)) continue;
            const int a2 = p[2 * st + 2], a10 = p[-2 * st - 2], a6 = p[-2 * st + 2], a14 = p[2 * st - 2];
            br = br && (a2 > hi || a10 > hi) && (a6 > hi || a14 > hi);
            dk = dk && (a2 < lo || a10 < lo) && (a6 < lo || a14 < lo);
            if (!(br || dk)) continue;
            int m = fast_arc_measure(p, step);
            if (m > threshold) score[(size_t)y * cols + x] = (uint8_t)(m - 1);
        }
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            const uint8_t* s = &score[(size_t)y * cols + x];
            int c = s[0];
            if (!c) continue;
            if (c > s[-1] && c > s[1] && c > s[-cols - 1] && c > s[-cols] && c > s[-cols + 1] &&
                c > s[cols - 1] && c > s[cols] && c > s[cols + 1])
                out.push_back({x, y, c});
        }
}

// ---------------------------------------------------------------------------------------
// cv::GaussianBlur(u8, Size(13,13), 2, 2, BORDER_REFLECT_101): OpenCV's fixed-point 8-bit
// path (imgproc/smooth.dispatch.cpp, bit-exact Q8 kernel).  Call site: ORBextractor.cpp:1092.
// ---------------------------------------------------------------------------------------
const int kGauss13[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};
inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}
void gaussian13_s2(const uint8_t* src, int rows, int cols, size_t sstep, uint8_t* dst, size_t dstep) {
    std::vector<uint16_t> H((size_t)rows * cols);
    for (int y = 0; y < rows; y++) {
        const uint8_t* S = src + (size_t)y * sstep;
        for (int x = 0; x < cols; x++) {
            int acc = 0;
            for (int i = 0; i < 13; i++) acc += kGauss13[i] * S[reflect101(x + i - 6, cols)];
            H[(size_t)y * cols + x] = (uint16_t)acc;  // <= 255*256
        }
    }
    for (int y = 0; y < rows; y++) {
        uint8_t* D = dst + (size_t)y * dstep;
        for (int x = 0; x < cols; x++) {
            uint32_t acc = 0;
            for (int i = 0; i < 13; i++) acc += (uint32_t)kGauss13[i] * H[(size_t)reflect101(y + i - 6, rows) * cols + x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------
// cv::fastAtan2 (core/mathfuncs_core.simd.hpp atan_f32), float, degrees.  Call site :103.
// Built with -ffp-contract=off so that no FMA is formed.
// ---------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------
// cosf / sinf as the reference gets them from its C library (ORBextractor.cpp:113, `using namespace std`
// -> std::cos(float) -> cosf).  Third-party dependency: GNU libc, 2.39 in this image (Ubuntu GLIBC
// 2.39-0ubuntu8.5); its float sine / cosine is the published algorithm of the Arm Optimized Routines
// (glibc sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h, sincosf_data.c): argument in double, quadrant
// n = round(x * 2/pi) by a scaled integer conversion, r = x - n * pi/2, a degree-7 sine or degree-8 cosine
// polynomial in double, ONE rounding to float at the end.  Restated here for |x| < 120 (the keypoint angle is
// fastAtan2 degrees times pi/180, so x is in [0, 2*pi]).  The float result does not depend on whether the
// double multiply-adds are fused (both forms were scanned against libm over the whole domain).
// ---------------------------------------------------------------------------------------
inline uint32_t f32_top12(float x) { uint32_t u; std::memcpy(&u, &x, 4); return (u >> 20) & 0x7ff; }
inline float sincosf_poly(double x, double x2, int tab, int n) {
    static const double kC[2][5] = {
        {0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16},
        {-0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16}};
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = x * x2, t = s2 + x2 * s3, x7 = x3 * x2, s = x + x3 * s1;
        return (float)(s + x7 * t);
    }
    const double* c = kC[tab];
    const double x4 = x2 * x2, c2 = c[3] + x2 * c[4], c1 = c[0] + x2 * c[1], x6 = x4 * x2, cc = c1 + x4 * c[2];
    return (float)(cc + x6 * c2);
}
void libm_sincosf(float y, float* sn, float* cs) {
    const double kSign[4] = {1.0, -1.0, -1.0, 1.0};
    double x = y;
    if (f32_top12(y) < 0x3f4) {                       // |y| < 0.75 (compared on the top 12 bits, as glibc does)
        const double x2 = x * x;
        if (f32_top12(y) < 0x398) { *sn = y; *cs = 1.0f; return; }   // |y| < 2^-12
        *sn = sincosf_poly(x, x2, 0, 0);
        *cs = sincosf_poly(x, x2, 0, 1);
        return;
    }
    const double r = x * 0x1.45F306DC9C883p+23;       // 2/pi * 2^24
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * 0x1.921FB54442D18p0;
    const double s = kSign[n & 3];
    const int tab = (n & 2) ? 1 : 0;
    *sn = sincosf_poly(x * s, x * x, tab, n);
    *cs = sincosf_poly(x * s, x * x, tab, n ^ 1);
}

// ---------------------------------------------------------------------------------------
// The extractor (ORBextractor.cpp:410-470 ctor, :1049-1140 operator()/ComputePyramid,
// :765-853 ComputeKeyPointsOctTree, :539-763 DistributeOctTree, :481-537 DivideNode).
// ---------------------------------------------------------------------------------------
struct Kp { float x, y, size, angle, response; int octave, class_id; };

struct Node {  // ExtractorNode, ORBextractor.h:32-43
    std::vector<Kp> keys;
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool noMore = false;
    long seq = 0;  // B2: creation sequence number replaces the heap address in the sort
};

void divide_node(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {  // :481-537
    const int halfX = (int)std::ceil((float)(n.URx - n.ULx) / 2);
    const int halfY = (int)std::ceil((float)(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (const Kp& kp : n.keys) {
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.keys.push_back(kp); else n3.keys.push_back(kp);
        } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
        else n4.keys.push_back(kp);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
}

struct Extractor {
    int nfeatures, nlevels, iniThFAST, minThFAST;
    double scaleFactor;  // the member is a double holding the float argument (ORBextractor.h:98)
    std::vector<float> scale, invScale;
    std::vector<int> featuresPerLevel, umax;
    // introspection (filled by run())
    std::vector<Image> pyramid;
    std::vector<std::vector<Kp>> candidates;   // per level, coords relative to (16,16)
    std::vector<std::vector<Kp>> levelKeys;    // per level after distribution+orientation (level coords)

    Extractor(int nf, float sf, int nl, int ini, int mn)
        : nfeatures(nf), nlevels(nl), iniThFAST(ini), minThFAST(mn), scaleFactor(sf) {
        scale.resize(nl); invScale.resize(nl);
        scale[0] = 1.0f;
        for (int i = 1; i < nl; i++) scale[i] = (float)(scale[i - 1] * scaleFactor);  // :421 (float*double)
        for (int i = 0; i < nl; i++) invScale[i] = 1.0f / scale[i];                    // :429
        featuresPerLevel.resize(nl);
        float factor = (float)(1.0f / scaleFactor);                                    // :436
        float nDesired = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl)));  // :437
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) {
            featuresPerLevel[l] = cvRoundF(nDesired);
            sum += featuresPerLevel[l];
            nDesired *= factor;
        }
        featuresPerLevel[nl - 1] = std::max(nfeatures - sum, 0);
        umax.resize(HALF_PATCH_SIZE + 1);                                              // :454-469
        int v, v0, vmax = cvFloorF(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
        int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cvRoundD(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    void level_size(int rows, int cols, int l, int& r, int& c) const {  // :1119-1120
        float s = invScale[l];
        c = cvRoundF((float)cols * s);
        r = cvRoundF((float)rows * s);
    }

    void compute_pyramid(const uint8_t* img, int rows, int cols, size_t step) {  // :1115-1140
        pyramid.assign(nlevels, Image());
        for (int l = 0; l < nlevels; l++) {
            int r, c;
            level_size(rows, cols, l, r, c);
            pyramid[l] = Image(r, c);
            if (l == 0)
                for (int y = 0; y < rows; y++) std::memcpy(pyramid[0].row(y), img + (size_t)y * step, cols);
            else
                resize_linear_u8(pyramid[l - 1].d.data(), pyramid[l - 1].rows, pyramid[l - 1].cols,
                                 pyramid[l - 1].cols, pyramid[l].d.data(), r, c, c);
            // the 19-px REFLECT_101 border (:1130,:1135) is never read downstream -> not materialised
        }
    }

    std::vector<Kp> distribute(const std::vector<Kp>& in, int minX, int maxX, int minY, int maxY, int N) {
        int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));  // :543
        if (nIni < 1) nIni = 1;                                           // B1
        const float hX = (float)(maxX - minX) / nIni;                     // :545
        std::list<Node> nodes;
        std::vector<Node*> ini(nIni);
        long seq = 0;
        for (int i = 0; i < nIni; i++) {                                  // :552-563
            Node ni;
            ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
            ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
            ni.BLx = ni.ULx; ni.BLy = maxY - minY;
            ni.BRx = ni.URx; ni.BRy = maxY - minY;
            ni.seq = seq++;
            nodes.push_back(ni);
            ini[i] = &nodes.back();
        }
        for (const Kp& kp : in) ini[(size_t)(kp.x / hX)]->keys.push_back(kp);  // :566-570
        for (auto lit = nodes.begin(); lit != nodes.end();) {                  // :572-585
            if (lit->keys.size() == 1) { lit->noMore = true; ++lit; }
            else if (lit->keys.empty()) lit = nodes.erase(lit);
            else ++lit;
        }
        bool finish = false;
        typedef std::pair<int, Node*> SP;
        std::vector<SP> sizePtr;
        auto push_children = [&](Node* kids[4], std::vector<SP>& vec, int* nToExpand) {
            for (int c = 0; c < 4; c++) {
                Node& n = *kids[c];
                if (n.keys.size() > 0) {
                    n.seq = seq++;
                    nodes.push_front(n);
                    if (n.keys.size() > 1) {
                        if (nToExpand) (*nToExpand)++;
                        vec.push_back(std::make_pair((int)n.keys.size(), &nodes.front()));
                        nodes.front().lit = nodes.begin();
                    }
                }
            }
        };
        while (!finish) {                                                      // :594-739
            int prevSize = (int)nodes.size();
            auto lit = nodes.begin();
            int nToExpand = 0;
            sizePtr.clear();
            while (lit != nodes.end()) {
                if (lit->noMore) { ++lit; continue; }
                Node n1, n2, n3, n4;
                divide_node(*lit, n1, n2, n3, n4);
                Node* kids[4] = {&n1, &n2, &n3, &n4};
                push_children(kids, sizePtr, &nToExpand);
                lit = nodes.erase(lit);
            }
            if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
                finish = true;
            } else if (((int)nodes.size() + nToExpand * 3) > N) {               // :673
                while (!finish) {
                    prevSize = (int)nodes.size();
                    std::vector<SP> prev = sizePtr;
                    sizePtr.clear();
                    std::sort(prev.begin(), prev.end(), [](const SP& a, const SP& b) {  // :684 + B2
                        if (a.first != b.first) return a.first < b.first;
                        return a.second->seq < b.second->seq;
                    });
                    for (int j = (int)prev.size() - 1; j >= 0; j--) {
                        Node n1, n2, n3, n4;
                        divide_node(*prev[j].second, n1, n2, n3, n4);
                        Node* kids[4] = {&n1, &n2, &n3, &n4};
                        push_children(kids, sizePtr, nullptr);
                        nodes.erase(prev[j].second->lit);
                        if ((int)nodes.size() >= N) break;
                    }
                    if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
                }
            }
        }
        std::vector<Kp> out;                                                   // :742-760
        out.reserve(nodes.size());
        for (Node& n : nodes) {
            const Kp* best = &n.keys[0];
            float maxResponse = best->response;
            for (size_t k = 1; k < n.keys.size(); k++)
                if (n.keys[k].response > maxResponse) { best = &n.keys[k]; maxResponse = n.keys[k].response; }
            out.push_back(*best);
        }
        return out;
    }

    float ic_angle(const Image& im, float px, float py) const {              // :77-104
        int m_01 = 0, m_10 = 0;
        const int step = im.cols;
        const uint8_t* center = im.row(cvRoundF(py)) + cvRoundF(px);
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int v_sum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int vp = center[u + v * step], vm = center[u - v * step];
                v_sum += (vp - vm);
                m_10 += u * (vp + vm);
            }
            m_01 += v * v_sum;
        }
        return fast_atan2((float)m_01, (float)m_10);
    }

    void compute_keypoints() {                                               // :765-853
        candidates.assign(nlevels, {});
        levelKeys.assign(nlevels, {});
        const float W = 30;
        for (int level = 0; level < nlevels; ++level) {
            const Image& im = pyramid[level];
            const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
            const int maxBorderX = im.cols - EDGE_THRESHOLD + 3, maxBorderY = im.rows - EDGE_THRESHOLD + 3;
            std::vector<Kp>& cand = candidates[level];
            const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            // nCols==0 / nRows==0: the reference evaluates ceil(x/0) but never uses it (loops are empty)
            const int wCell = nCols > 0 ? (int)std::ceil(width / nCols) : 0;
            const int hCell = nRows > 0 ? (int)std::ceil(height / nRows) : 0;
            std::vector<RawKp> cell;
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBorderY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBorderY - 3) continue;
                if (maxY > maxBorderY) maxY = (float)maxBorderY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBorderX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBorderX - 6) continue;
                    if (maxX > maxBorderX) maxX = (float)maxBorderX;
                    const int y0 = (int)iniY, y1 = (int)maxY, x0 = (int)iniX, x1 = (int)maxX;
                    const uint8_t* roi = im.row(y0) + x0;
                    fast9_16(roi, y1 - y0, x1 - x0, im.cols, iniThFAST, cell);
                    if (cell.empty()) fast9_16(roi, y1 - y0, x1 - x0, im.cols, minThFAST, cell);
                    for (const RawKp& r : cell)
                        cand.push_back({(float)r.x + j * wCell, (float)r.y + i * hCell, 7.f, -1.f, (float)r.score, 0, -1});
                }
            }
            std::vector<Kp> keys = distribute(cand, minBorderX, maxBorderX, minBorderY, maxBorderY, featuresPerLevel[level]);
            const int scaledPatchSize = (int)(PATCH_SIZE * scale[level]);   // :837
            for (Kp& k : keys) {
                k.x += minBorderX; k.y += minBorderY; k.octave = level; k.size = (float)scaledPatchSize;
            }
            levelKeys[level] = keys;
        }
        for (int level = 0; level < nlevels; ++level)
            for (Kp& k : levelKeys[level]) k.angle = ic_angle(pyramid[level], k.x, k.y);
    }

    static void orb_descriptor(const Kp& kpt, const Image& img, uint8_t* desc) {  // :108-147
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        float angle = (float)kpt.angle * factorPI;
        float a, b;
        libm_sincosf(angle, &b, &a);                                                    // A5: cosf / sinf of :113
        const int step = img.cols;
        const uint8_t* center = img.row(cvRoundF(kpt.y)) + cvRoundF(kpt.x);
        const signed char* pat = kPattern;
        auto val = [&](int idx) -> int {
            float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
            float fy = px * b + py * a, fx = px * a - py * b;   // no FMA (-ffp-contract=off)
            return center[cvRoundF(fy) * step + cvRoundF(fx)];
        };
        for (int i = 0; i < 32; ++i, pat += 32) {
            int v = 0;
            for (int k = 0; k < 8; k++) {
                int t0 = val(2 * k), t1 = val(2 * k + 1);
                v |= (t0 < t1) << k;
            }
            desc[i] = (uint8_t)v;
        }
    }

    // operator(): returns keypoints (image coords) and n x 32 descriptors.
    void run(const uint8_t* img, int rows, int cols, size_t step, std::vector<Kp>& kps, std::vector<uint8_t>& desc) {
        kps.clear(); desc.clear();
        compute_pyramid(img, rows, cols, step);
        compute_keypoints();
        for (int level = 0; level < nlevels; ++level) {                      // :1082-1112
            std::vector<Kp>& keys = levelKeys[level];
            if (keys.empty()) continue;
            Image work(pyramid[level].rows, pyramid[level].cols);
            gaussian13_s2(pyramid[level].d.data(), work.rows, work.cols, work.cols, work.d.data(), work.cols);
            size_t off = desc.size();
            desc.resize(off + keys.size() * 32, 0);
            for (size_t i = 0; i < keys.size(); i++) orb_descriptor(keys[i], work, &desc[off + i * 32]);
            for (const Kp& k0 : keys) {
                Kp k = k0;
                if (level != 0) { k.x *= scale[level]; k.y *= scale[level]; }
                kps.push_back(k);
            }
        }
    }
};

void copy_kps(const std::vector<Kp>& v, orc_keypoint* out) {
    static_assert(sizeof(orc_keypoint) == 28, "cv::KeyPoint layout");
    for (size_t i = 0; i < v.size(); i++) {
        out[i].x = v[i].x; out[i].y = v[i].y; out[i].size = v[i].size; out[i].angle = v[i].angle;
        out[i].response = v[i].response; out[i].octave = v[i].octave; out[i].class_id = v[i].class_id;
    }
}

}  // namespace

extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int srows, int scols, int sstep, uint8_t* dst, int drows,
                          int dcols, int dstep) {
    resize_linear_u8(src, srows, scols, sstep, dst, drows, dcols, dstep);
}

int orc_fast9_16(const uint8_t* img, int rows, int cols, int step, int threshold, int* xys, int cap) {
    std::vector<RawKp> v;
    fast9_16(img, rows, cols, step, threshold, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) {
        xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score;
    }
    return (int)v.size();
}

void orc_gaussian13_s2(const uint8_t* src, int rows, int cols, int sstep, uint8_t* dst, int dstep) {
    gaussian13_s2(src, rows, cols, sstep, dst, dstep);
}

float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_sincosf(float x, float* s, float* c) { libm_sincosf(x, s, c); }
// number of floats with bit patterns first .. first+count-1 on which libm_sincosf differs from THIS host's cosf / sinf
long orc_sincosf_scan(uint32_t first, uint32_t count) {
    long bad = 0;
    for (uint32_t i = 0; i < count; i++) {
        uint32_t u = first + i;
        float x, s, c;
        std::memcpy(&x, &u, 4);
        libm_sincosf(x, &s, &c);
        const float hs = sinf(x), hc = cosf(x);
        bad += (std::memcmp(&s, &hs, 4) != 0) + (std::memcmp(&c, &hc, 4) != 0);
    }
    return bad;
}

void orc_pattern(signed char* out1024) { std::memcpy(out1024, kPattern, 1024); }

void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    return new Extractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }

void orc_extractor_tables(void* h, float* scale, float* inv_scale, int* features_per_level, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; i++) {
        scale[i] = e->scale[i]; inv_scale[i] = e->invScale[i]; features_per_level[i] = e->featuresPerLevel[i];
    }
    for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

int orc_extractor_run(void* h, const uint8_t* img, int rows, int cols, int step, orc_keypoint* kps,
                      uint8_t* desc, int cap) {
    Extractor* e = (Extractor*)h;
    std::vector<Kp> k; std::vector<uint8_t> d;
    e->run(img, rows, cols, step, k, d);
    int n = (int)k.size();
    if (n <= cap) { copy_kps(k, kps); std::memcpy(desc, d.data(), d.size()); }
    return n;
}

void orc_extractor_level_size(void* h, int rows, int cols, int level, int* lrows, int* lcols) {
    ((Extractor*)h)->level_size(rows, cols, level, *lrows, *lcols);
}
// Introspection of the last run().
void orc_extractor_level_image(void* h, int level, uint8_t* out) {
    Extractor* e = (Extractor*)h;
    std::memcpy(out, e->pyramid[level].d.data(), e->pyramid[level].d.size());
}
int orc_extractor_candidates(void* h, int level, int* xys, int cap) {
    Extractor* e = (Extractor*)h;
    const std::vector<Kp>& c = e->candidates[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) {
        xys[3 * i] = (int)c[i].x; xys[3 * i + 1] = (int)c[i].y; xys[3 * i + 2] = (int)c[i].response;
    }
    return (int)c.size();
}
int orc_extractor_level_keys(void* h, int level, orc_keypoint* out, int cap) {
    Extractor* e = (Extractor*)h;
    if ((int)e->levelKeys[level].size() <= cap) copy_kps(e->levelKeys[level], out);
    return (int)e->levelKeys[level].size();
}
// Stand-alone DistributeOctTree on a caller-supplied candidate list (x,y,response triples,
// coordinates relative to (minX,minY)); returns the selected keys in list order.
int orc_distribute(const int* xys, int n, int minX, int maxX, int minY, int maxY, int N, int* out_xys) {
    Extractor e(2000, 1.2f, 6, 12, 7);
    std::vector<Kp> in(n);
    for (int i = 0; i < n; i++) in[i] = {(float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1.f, (float)xys[3 * i + 2], 0, -1};
    std::vector<Kp> out = e.distribute(in, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < out.size(); i++) {
        out_xys[3 * i] = (int)out[i].x; out_xys[3 * i + 1] = (int)out[i].y; out_xys[3 * i + 2] = (int)out[i].response;
    }
    return (int)out.size();
}

}  // extern "C"
