// TEST INFRASTRUCTURE ONLY -- a stand-in for the slice of the OpenCV C++ API that the reference's hot-path
// sources use, written for this repository so that
//     /root/reference/thirdparty/ORBextractor.cpp
//     /root/reference/src/core/FEAmatcher.cpp
//     /root/reference/src/core/frame.cpp
//     /root/reference/src/util/util.cpp:1-43  (Util::ComputeIntersection)
// compile UNMODIFIED, where they lie, into oracle/_ref/ (recipe: oracle/build_ref.sh).  OpenCV's C++ headers are
// not installed in this image.  Nothing in the product (diasss_b200/, include/) includes this file.
//
// What is behind the names:
//   * containers and plumbing (Mat with shared buffers and ROI views, MatExpr-style zeros()/ones() that assign IN
//     PLACE into an equally sized header as OpenCV's do -- ORBextractor.cpp:1037 relies on it --, InputArray /
//     OutputArray proxies, Mat_<T> with the comma initialiser, KeyPoint, Point_, Size, Rect, Scalar, RNG) are
//     re-implemented here with the documented OpenCV semantics;
//   * the numeric primitives (resize INTER_LINEAR 8U, FAST 9/16 + NMS, GaussianBlur 13x13 sigma 2, fastAtan2, mean)
//     forward to the oracle's primitives, which are pinned bit-exactly against the real OpenCV 4.13 (cv2) by
//     tests/test_oracle_primitives.py and tests/golden/primitives_cv2.npz (ref_cv_impl.cpp);
//   * everything the reference only references from dead or switched-off code (SIFT, drawMatches, imshow, norm,
//     KeyPointsFilter) exists so that the sources compile; calling SIFT aborts.
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <list>
#include <map>
#include <utility>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef uint64_t uint64;

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum BorderTypes {
    BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
    BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16
};
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
enum WindowFlags { WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1 };
enum struct DrawMatchesFlags { DEFAULT = 0, DRAW_OVER_OUTIMG = 1, NOT_DRAW_SINGLE_POINTS = 2, DRAW_RICH_KEYPOINTS = 4 };

// cvRound = cvtss2si / cvtsd2si (round half to even in the default rounding mode); cvFloor / cvCeil as in
// core/fast_math.hpp.
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (v < i); }
inline int cvFloor(float v) { int i = (int)v; return i - (v < i); }
inline int cvFloor(int v) { return v; }
inline int cvCeil(double v) { int i = (int)v; return i + (v > i); }
inline int cvCeil(float v) { int i = (int)v; return i + (v > i); }
inline int cvCeil(int v) { return v; }

template <typename T> inline T saturate_cast(double v);
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline float saturate_cast<float>(double v) { return (float)v; }
template <> inline double saturate_cast<double>(double v) { return v; }

float fastAtan2(float y, float x);   // ref_cv_impl.cpp -> the oracle's cv2-pinned polynomial

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) {   // core/types.hpp
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}
template <> inline Point_<int>& operator*=(Point_<int>& a, float b) {
    a.x = cvRound(a.x * b);
    a.y = cvRound(a.y * b);
    return a;
}
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T>
struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar_<T> all(T v) { return Scalar_<T>(v, v, v, v); }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

struct KeyPoint {   // 28 bytes, the layout the C ABI's dsx_keypoint mirrors
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

inline size_t stub_elem_size(int type) {
    switch (type & 7) {
        case CV_8U: case CV_8S: return 1;
        case CV_16U: case CV_16S: return 2;
        case CV_32S: case CV_32F: return 4;
        default: return 8;
    }
}

struct MatInit { int rows, cols, type; double value; };   // what Mat::zeros / Mat::ones return (MatExpr)

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;

class Mat {
public:
    int flags, rows, cols;
    uchar* data;
    size_t step;

    Mat() : flags(0), rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); fill(s[0]); }
    Mat(Size sz, int type, const Scalar& s) : Mat() { create(sz.height, sz.width, type); fill(s[0]); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : flags(type), rows(r), cols(c), data((uchar*)ext) {
        step = step_ ? step_ : (size_t)c * stub_elem_size(type);
    }
    Mat(const MatInit& e) : Mat() { *this = e; }

    // MatExpr assignment: create() (a no-op for an equally sized and typed header, so a ROI view is written IN
    // PLACE: core/matrix_expressions.cpp MatOp_Initializer::assign) followed by a fill.
    Mat& operator=(const MatInit& e) {
        create(e.rows, e.cols, e.type);
        fill(e.value);
        return *this;
    }

    static MatInit zeros(int r, int c, int type) { return MatInit{r, c, type, 0.0}; }
    static MatInit zeros(Size sz, int type) { return MatInit{sz.height, sz.width, type, 0.0}; }
    static MatInit ones(int r, int c, int type) { return MatInit{r, c, type, 1.0}; }
    static MatInit ones(Size sz, int type) { return MatInit{sz.height, sz.width, type, 1.0}; }

    void create(int r, int c, int type) {   // Mat::create: keeps the buffer when shape and type already agree
        if (data && r == rows && c == cols && type == this->type()) return;
        release();
        rows = r; cols = c; flags = type;
        step = (size_t)c * stub_elem_size(type);
        if (r <= 0 || c <= 0) return;
        // slack behind the last row: Frame::GetFilteredMask stamps up to 5 rows / 5 bytes past the image
        // (frame.cpp:100-102); the stand-in keeps those writes inside its own allocation
        const size_t bytes = (size_t)(r + 8) * step + 64;
        uchar* p = (uchar*)std::calloc(bytes, 1);
        if (!p) std::abort();
        buf_ = std::shared_ptr<uchar>(p, [](uchar* q) { std::free(q); });
        data = p;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    int type() const { return flags & 7; }
    int depth() const { return flags & 7; }
    int channels() const { return 1; }
    size_t elemSize() const { return stub_elem_size(type()); }
    size_t elemSize1() const { return stub_elem_size(type()); }
    size_t step1() const { return step / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }

    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }

    Mat rowRange(int r0, int r1) const { Mat m(*this); m.data = data + (size_t)r0 * step; m.rows = r1 - r0; return m; }
    Mat colRange(int c0, int c1) const { Mat m(*this); m.data = data + (size_t)c0 * elemSize(); m.cols = c1 - c0; return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }

    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type());
        const size_t rb = (size_t)cols * elemSize();
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, rb);
        return m;
    }
    void copyTo(Mat& dst) const {
        dst.create(rows, cols, type());
        const size_t rb = (size_t)cols * elemSize();
        for (int r = 0; r < rows; r++) std::memmove(dst.data + (size_t)r * dst.step, data + (size_t)r * step, rb);
    }
    void convertTo(Mat& dst, int rtype) const;          // ref_cv_impl.cpp
    void convertTo(OutputArray dst, int rtype) const;

    // Mat::push_back(const Mat&): appends the rows of m; an empty matrix takes m's shape and type
    void push_back(const Mat& m) {
        if (m.empty()) return;
        const int nc = empty() ? m.cols : cols, ty = empty() ? m.type() : type();
        Mat nm(rows + m.rows, nc, ty);
        const size_t rb = (size_t)nc * stub_elem_size(ty);
        for (int r = 0; r < rows; r++) std::memcpy(nm.data + (size_t)r * nm.step, data + (size_t)r * step, rb);
        for (int r = 0; r < m.rows; r++) std::memcpy(nm.data + (size_t)(rows + r) * nm.step, m.data + (size_t)r * m.step, rb);
        *this = nm;
    }

    double get(int r, int c) const {
        const uchar* p = data + (size_t)r * step + (size_t)c * elemSize();
        switch (type()) {
            case CV_8U: return *p;
            case CV_8S: return *(const signed char*)p;
            case CV_16U: return *(const ushort*)p;
            case CV_16S: return *(const short*)p;
            case CV_32S: return *(const int*)p;
            case CV_32F: return *(const float*)p;
            default: return *(const double*)p;
        }
    }
    void fill(double v) {
        for (int r = 0; r < rows; r++) {
            uchar* p = data + (size_t)r * step;
            switch (type()) {
                case CV_8U: case CV_8S: std::memset(p, (int)saturate_cast<uchar>(v), (size_t)cols); break;
                case CV_16U: case CV_16S: for (int c = 0; c < cols; c++) ((short*)p)[c] = (short)cvRound(v); break;
                case CV_32S: for (int c = 0; c < cols; c++) ((int*)p)[c] = cvRound(v); break;
                case CV_32F: for (int c = 0; c < cols; c++) ((float*)p)[c] = (float)v; break;
                default: for (int c = 0; c < cols; c++) ((double*)p)[c] = v; break;
            }
        }
    }

protected:
    std::shared_ptr<uchar> buf_;
};

template <typename T> struct stub_type_of;
template <> struct stub_type_of<uchar> { enum { value = CV_8U }; };
template <> struct stub_type_of<int> { enum { value = CV_32S }; };
template <> struct stub_type_of<float> { enum { value = CV_32F }; };
template <> struct stub_type_of<double> { enum { value = CV_64F }; };

template <typename T>
class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, stub_type_of<T>::value) {}
    Mat_(const Mat& m) : Mat(m) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
};
typedef Mat_<uchar> Mat1b;
typedef Mat_<int> Mat1i;
typedef Mat_<float> Mat1f;
typedef Mat_<double> Mat1d;

// (Mat_<T>(r, c) << v0, v1, ...): row-major fill, converted back to Mat_<T> (core/mat.inl.hpp MatCommaInitializer_)
template <typename T>
class MatCommaInitializer_ {
public:
    explicit MatCommaInitializer_(const Mat_<T>& m) : m_(m), i_(0) {}
    template <typename T2> MatCommaInitializer_<T>& operator,(T2 v) {
        m_(i_ / m_.cols, i_ % m_.cols) = (T)v;
        i_++;
        return *this;
    }
    operator Mat_<T>() const { return m_; }
private:
    Mat_<T> m_;
    int i_;
};
template <typename T, typename T2>
inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 v) {
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}

// InputArray / OutputArray: references to proxy objects, as in OpenCV; only the Mat flavour is needed here.
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
    int type() const { return m_ ? m_->type() : 0; }
    Size size() const { return m_ ? m_->size() : Size(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size sz, int type) const { m_->create(sz.height, sz.width, type); }
    void release() const { m_->release(); }
    Mat& getMatRef() const { return *m_; }
    bool needed() const { return m_ != nullptr; }
};
inline InputArray noArray() { static _InputArray none; return none; }

// ---- numeric primitives (ref_cv_impl.cpp) -------------------------------------------------------------------------
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0,
                  int borderType = BORDER_DEFAULT);
void minMaxLoc(InputArray src, double* minVal, double* maxVal = nullptr, Point* minLoc = nullptr, Point* maxLoc = nullptr);
Scalar mean(InputArray src);
double norm(InputArray a, InputArray b, int normType = NORM_L2);

// cv::RNG: multiply-with-carry generator (core/operations.hpp); uniform(int,int) = a + next() % (b - a)
class RNG {
public:
    uint64 state;
    RNG() : state(0xffffffff) {}
    RNG(uint64 s) : state(s ? s : 0xffffffff) {}
    unsigned next() {
        state = (uint64)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    operator unsigned() { return next(); }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (b - a) + a); }
    float uniform(float a, float b) { return ((float)next()) * 2.3283064365386962890625e-10f * (b - a) + a; }
    double uniform(double a, double b) {
        unsigned t = next();
        return (double)(((uint64)t << 32) | next()) * 5.4210108624275221700372640043497e-20 * (b - a) + a;
    }
};
RNG& theRNG();
inline void setRNGSeed(int seed) { theRNG() = RNG((uint64)(unsigned)seed); }

// ---- names that only dead / switched-off reference code touches -----------------------------------------------------
template <typename T> using Ptr = std::shared_ptr<T>;

class Feature2D {
public:
    virtual ~Feature2D() {}
    virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors);
};
class SIFT : public Feature2D {
public:
    static Ptr<SIFT> create() { return std::make_shared<SIFT>(); }
};
typedef SIFT SiftDescriptorExtractor;
typedef SIFT SiftFeatureDetector;

class KeyPointsFilter {
public:
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

void drawMatches(InputArray img1, const std::vector<KeyPoint>& k1, InputArray img2, const std::vector<KeyPoint>& k2,
                 const std::vector<DMatch>& matches, InputOutputArray out, const Scalar& matchColor = Scalar::all(-1),
                 const Scalar& pointColor = Scalar::all(-1), const std::vector<char>& mask = std::vector<char>(),
                 DrawMatchesFlags flags = DrawMatchesFlags::DEFAULT);
void drawKeypoints(InputArray image, const std::vector<KeyPoint>& kps, InputOutputArray out,
                   const Scalar& color = Scalar::all(-1), DrawMatchesFlags flags = DrawMatchesFlags::DEFAULT);
void namedWindow(const std::string& name, int flags = WINDOW_AUTOSIZE);
void imshow(const std::string& name, InputArray img);
int waitKey(int delay = 0);

// ---- test probe: every cv::FAST call can be recorded (the reference keeps its per-level candidate lists in
// locals, ORBextractor.cpp:782; the log makes them observable without touching the source) ---------------------------
namespace stub {
struct FastCall {
    const uchar* data;
    size_t step;
    int rows, cols, threshold;
    std::vector<KeyPoint> out;
};
extern std::vector<FastCall>* fast_log;   // null = off
extern int mean_order;                    // 0 = sequential sum, 1 = the library's 32-lane order (orc_mean)
}  // namespace stub

}  // namespace cv
