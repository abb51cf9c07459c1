// TEST INFRASTRUCTURE ONLY -- the context Optimizer::GetKpsPairs (src/core/optimizer.cpp:575-639) needs to compile on
// its own: the reference's optimizer.h pulls in GTSAM, which is not installed here.  oracle/build_ref.sh prepends
// this header to the function's own lines, taken from the reference file by sed, and closes the namespace.
// gtsam::Vector7 = Eigen::Matrix<double, 7, 1>: only its comma initialiser ((Vector7() << a, b, ...).finished()) is used.
#pragma once
#include <bits/stdc++.h>
#include <opencv2/opencv.hpp>

namespace gtsam {
struct Vector7 {
    double v[7];
    int n_;
    Vector7() : n_(0) { for (double& x : v) x = 0; }
    template <typename T> Vector7& operator<<(T x) { v[n_++] = (double)x; return *this; }
    template <typename T> Vector7& operator,(T x) { v[n_++] = (double)x; return *this; }
    Vector7& finished() { return *this; }
    double operator()(int i) const { return v[i]; }
};
}  // namespace gtsam

namespace Diasss
{
using namespace std;
using namespace cv;
using namespace gtsam;

class Optimizer
{
public:
    std::vector<Vector7> static GetKpsPairs(const bool &USE_ANNO, const cv::Mat &kps, const int &id_s, const int &id_t,
                                            const std::vector<double> &alts_s, const std::vector<double> &gras_s,
                                            const std::vector<double> &alts_t, const std::vector<double> &gras_t);
};
