// TEST INFRASTRUCTURE ONLY -- the non-inline half of the OpenCV stand-in (opencv2/opencv.hpp in this directory).
// The numeric primitives forward to the oracle's implementations (oracle/orb_oracle.cpp, match_oracle.cpp), which
// are pinned bit-exactly against the real OpenCV 4.13: the reference sources compiled against this stand-in thus
// see the arithmetic a real OpenCV would give them, and everything above the primitives is the reference's own code.
#include <opencv2/opencv.hpp>

#include <cstdio>

#include "../oracle_capi.h"

namespace cv {

namespace stub {
std::vector<FastCall>* fast_log = nullptr;
int mean_order = 1;
}  // namespace stub

static void stub_fail(const char* what) {
    std::fprintf(stderr, "oracle/_ref OpenCV stand-in: %s\n", what);
    std::abort();
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::resize: only the form the reference uses (8UC1, INTER_LINEAR, explicit dsize; ORBextractor.cpp:1128)
void resize(InputArray src_, OutputArray dst_, Size dsize, double, double, int interpolation) {
    Mat src = src_.getMat();
    if (src.type() != CV_8U || interpolation != INTER_LINEAR || dsize.width <= 0 || dsize.height <= 0)
        stub_fail("resize: only 8UC1 INTER_LINEAR with an explicit size is provided");
    dst_.create(dsize.height, dsize.width, CV_8U);
    Mat dst = dst_.getMat();
    orc_resize_linear_u8(src.data, src.rows, src.cols, (int)src.step, dst.data, dst.rows, dst.cols, (int)dst.step);
}

static inline int border_index(int p, int n, int type) {
    if ((unsigned)p < (unsigned)n) return p;
    if (n == 1) return 0;
    switch (type) {
        case BORDER_REPLICATE: return p < 0 ? 0 : n - 1;
        case BORDER_REFLECT: while ((unsigned)p >= (unsigned)n) p = p < 0 ? -p - 1 : 2 * n - 1 - p; return p;
        case BORDER_WRAP: p %= n; return p < 0 ? p + n : p;
        default: while ((unsigned)p >= (unsigned)n) p = p < 0 ? -p : 2 * n - 2 - p; return p;   // REFLECT_101
    }
}

// cv::copyMakeBorder, 8U/any depth, non-constant borders.  The reference's two calls (ORBextractor.cpp:1130,1135)
// have src = the centre ROI of dst (same buffer): the centre is moved first (a no-op then), the frame is filled from it.
void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value) {
    Mat src = src_.getMat();
    const int bt = borderType & ~BORDER_ISOLATED;
    const size_t es = src.elemSize();
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    for (int r = 0; r < src.rows; r++)
        std::memmove(dst.data + (size_t)(r + top) * dst.step + (size_t)left * es, src.data + (size_t)r * src.step,
                     (size_t)src.cols * es);
    if (bt == BORDER_CONSTANT) {
        for (int r = 0; r < dst.rows; r++)
            for (int c = 0; c < dst.cols; c++)
                if (r < top || r >= top + src.rows || c < left || c >= left + src.cols) {
                    Mat px = dst.rowRange(r, r + 1).colRange(c, c + 1);
                    px.fill(value[0]);
                }
        return;
    }
    for (int r = top; r < top + src.rows; r++) {   // left / right margins of the centre rows
        uchar* row = dst.data + (size_t)r * dst.step;
        for (int c = 0; c < dst.cols; c++) {
            if (c >= left && c < left + src.cols) continue;
            const int sc = border_index(c - left, src.cols, bt);
            std::memcpy(row + (size_t)c * es, row + (size_t)(sc + left) * es, es);
        }
    }
    for (int r = 0; r < dst.rows; r++) {           // top / bottom rows from the completed centre rows
        if (r >= top && r < top + src.rows) continue;
        const int sr = border_index(r - top, src.rows, bt);
        std::memcpy(dst.data + (size_t)r * dst.step, dst.data + (size_t)(sr + top) * dst.step, (size_t)dst.cols * es);
    }
}

// cv::FAST(image, keypoints, threshold, true) = FAST_t<16>; KeyPoint(x, y, 7.f, -1, score)
void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    Mat img = image_.getMat();
    if (img.type() != CV_8U || !nonmaxSuppression) stub_fail("FAST: only 8UC1 with non-max suppression is provided");
    keypoints.clear();
    const int cap = std::max(16, img.rows * img.cols / 4 + 16);
    std::vector<int> xys((size_t)cap * 3);
    const int n = orc_fast9_16(img.data, img.rows, img.cols, (int)img.step, threshold, xys.data(), cap);
    if (n > cap) stub_fail("FAST: more keypoints than a 3x3 non-max suppression allows");
    keypoints.reserve(n);
    for (int i = 0; i < n; i++)
        keypoints.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1.f, (float)xys[3 * i + 2]));
    if (stub::fast_log) stub::fast_log->push_back(stub::FastCall{img.data, img.step, img.rows, img.cols, threshold, keypoints});
}

// cv::GaussianBlur: only 8UC1, 13x13, sigma 2, REFLECT_101 (ORBextractor.cpp:1092); works in place
void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
    Mat src = src_.getMat();
    if (src.type() != CV_8U || ksize.width != 13 || ksize.height != 13 || sigmaX != 2 || (sigmaY != 2 && sigmaY != 0) ||
        (borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101)
        stub_fail("GaussianBlur: only 8UC1 13x13 sigma 2 REFLECT_101 is provided");
    dst_.create(src.rows, src.cols, CV_8U);
    Mat dst = dst_.getMat();
    orc_gaussian13_s2(src.data, src.rows, src.cols, (int)src.step, dst.data, (int)dst.step);
}

void minMaxLoc(InputArray src_, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc) {
    Mat m = src_.getMat();
    if (m.empty()) stub_fail("minMaxLoc: empty matrix");
    if (m.type() == CV_64F && !minLoc && !maxLoc) {   // the form the hot path uses (FEAmatcher.cpp:71-72, util.cpp:21-26)
        double lo[4], hi[4];
        for (int k = 0; k < 4; k++) lo[k] = hi[k] = m.at<double>(0, 0);
        for (int r = 0; r < m.rows; r++) {
            const double* p = m.ptr<double>(r);
            int c = 0;
            for (; c + 4 <= m.cols; c += 4)
                for (int k = 0; k < 4; k++) { lo[k] = p[c + k] < lo[k] ? p[c + k] : lo[k]; hi[k] = p[c + k] > hi[k] ? p[c + k] : hi[k]; }
            for (; c < m.cols; c++) { lo[0] = p[c] < lo[0] ? p[c] : lo[0]; hi[0] = p[c] > hi[0] ? p[c] : hi[0]; }
        }
        for (int k = 1; k < 4; k++) { lo[0] = lo[k] < lo[0] ? lo[k] : lo[0]; hi[0] = hi[k] > hi[0] ? hi[k] : hi[0]; }
        if (minVal) *minVal = lo[0];
        if (maxVal) *maxVal = hi[0];
        return;
    }
    double mn = m.get(0, 0), mx = mn;
    Point pmn(0, 0), pmx(0, 0);
    for (int r = 0; r < m.rows; r++)
        for (int c = 0; c < m.cols; c++) {
            const double v = m.get(r, c);
            if (v < mn) { mn = v; pmn = Point(c, r); }
            if (v > mx) { mx = v; pmx = Point(c, r); }
        }
    if (minVal) *minVal = mn;
    if (maxVal) *maxVal = mx;
    if (minLoc) *minLoc = pmn;
    if (maxLoc) *maxLoc = pmx;
}

// cv::mean: CV_64F goes through the oracle (summation order selected by stub::mean_order); other depths sequential
Scalar mean(InputArray src_) {
    Mat m = src_.getMat();
    if (m.empty()) return Scalar();
    if (m.type() == CV_64F) {
        Mat c = m.isContinuous() ? m : m.clone();
        return Scalar(orc_mean((const double*)c.data, c.rows, c.cols, stub::mean_order));
    }
    double s = 0;
    for (int r = 0; r < m.rows; r++)
        for (int c = 0; c < m.cols; c++) s += m.get(r, c);
    return Scalar(s / (double)m.total());
}

double norm(InputArray a_, InputArray b_, int normType) {
    Mat a = a_.getMat(), b = b_.getMat();
    double acc = 0;
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) {
            const double d = std::fabs(a.get(r, c) - b.get(r, c));
            if (normType == NORM_INF) acc = std::max(acc, d);
            else if (normType == NORM_L1) acc += d;
            else acc += d * d;
        }
    return normType == NORM_L2 ? std::sqrt(acc) : acc;
}

void Mat::convertTo(Mat& dst, int rtype) const {
    if (rtype < 0) rtype = type();
    Mat out(rows, cols, rtype);   // a fresh buffer also covers the in-place call of frame.cpp:78
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) {
            const double v = get(r, c);
            uchar* p = out.data + (size_t)r * out.step + (size_t)c * out.elemSize();
            switch (rtype) {
                case CV_8U: *p = saturate_cast<uchar>(v); break;
                case CV_32S: *(int*)p = saturate_cast<int>(v); break;
                case CV_32F: *(float*)p = (float)v; break;
                case CV_64F: *(double*)p = v; break;
                default: stub_fail("convertTo: depth not provided");
            }
        }
    dst = out;
}
void Mat::convertTo(OutputArray dst, int rtype) const { convertTo(dst.getMatRef(), rtype); }

RNG& theRNG() {
    static thread_local RNG rng;
    return rng;
}

void Feature2D::compute(InputArray, std::vector<KeyPoint>&, OutputArray) {
    stub_fail("SIFT::compute reached: the library must be built in ORB mode (switch S1)");
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& kps, int n) {   // only ComputeKeyPointsOld (dead) calls it
    if (n >= 0 && (int)kps.size() > n) {
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(),
                         [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        const float ambiguous = kps[n - 1].response;
        auto end = std::partition(kps.begin() + n, kps.end(), [=](const KeyPoint& k) { return k.response >= ambiguous; });
        kps.resize(end - kps.begin());
    }
}

void drawMatches(InputArray, const std::vector<KeyPoint>&, InputArray, const std::vector<KeyPoint>&,
                 const std::vector<DMatch>&, InputOutputArray, const Scalar&, const Scalar&, const std::vector<char>&,
                 DrawMatchesFlags) {}
void drawKeypoints(InputArray, const std::vector<KeyPoint>&, InputOutputArray, const Scalar&, DrawMatchesFlags) {}
void namedWindow(const std::string&, int) {}
void imshow(const std::string&, InputArray) {}
int waitKey(int) { return -1; }

}  // namespace cv
