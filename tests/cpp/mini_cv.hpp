// TEST INFRASTRUCTURE ONLY.  A minimal stand-in for the handful of OpenCV core types the diasss call surface
// uses (cv::Mat, cv::KeyPoint, cv::InputArray / cv::OutputArray), so that include/diasss_b200/shim.hpp -- the
// C++ drop-in for ORB_SLAM2::ORBextractor and Diasss::FEAmatcher -- can be compiled and exercised in an image
// without OpenCV's C++ headers.  With real OpenCV the shim includes <opencv2/core.hpp> instead
// and this file is unused.  Only the members the shim and the reference's call sites touch are provided.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

typedef unsigned char uchar;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

struct KeyPoint {   // 28 bytes, same layout as OpenCV's
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

inline size_t elem_size(int type) {
    switch (type) {
        case CV_8U: return 1;
        case CV_32S: case CV_32F: return 4;
        default: return 8;
    }
}

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), type_(type) {
        step = step_ ? step_ : (size_t)c * elem_size(type);
    }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && data) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elem_size(type);
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step + 16, (uchar)0);
        data = buf_->data();
    }
    void release() { rows = cols = 0; data = nullptr; step = 0; buf_.reset(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return elem_size(type_); }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    Mat row(int r) const {
        Mat m; m.rows = 1; m.cols = cols; m.type_ = type_; m.step = step; m.data = data + (size_t)r * step; m.buf_ = buf_;
        return m;
    }
    // append the rows of m (same cols / type); always re-packs into an own continuous buffer
    void push_back(const Mat& m) {
        if (m.empty()) return;
        const int nc = empty() ? m.cols : cols, ty = empty() ? m.type_ : type_;
        const size_t rb = (size_t)nc * elem_size(ty);
        auto nb = std::make_shared<std::vector<uchar>>((size_t)(rows + m.rows) * rb + 16, (uchar)0);
        for (int r = 0; r < rows; r++) std::memcpy(nb->data() + (size_t)r * rb, data + (size_t)r * step, rb);
        for (int r = 0; r < m.rows; r++) std::memcpy(nb->data() + (size_t)(rows + r) * rb, m.data + (size_t)r * m.step, rb);
        rows += m.rows; cols = nc; type_ = ty; step = rb; buf_ = nb; data = nb->data();
    }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, m.step);
        return m;
    }

private:
    int type_ = 0;
    std::shared_ptr<std::vector<uchar>> buf_;
};

// cv::InputArray / cv::OutputArray are references to proxy objects in OpenCV; a thin proxy over Mat suffices here.
class _InputArray {
public:
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return *m_; }
    bool empty() const { return m_->empty(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
    Mat& getMatRef() const { return *m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
