// TEST INFRASTRUCTURE ONLY -- declaration of the ONE Diasss::Util member that is compiled into oracle/_ref:
// Util::ComputeIntersection (the body is /root/reference/src/util/util.cpp:13-43, compiled from there by
// oracle/build_ref.sh).  The reference's own util.h (src/util/util.h:5-12) pulls in Boost.Filesystem and Eigen for
// LoadInputData and the viewers, which are off the hot path and not installed in this image.
#pragma once
#include <iostream>
#include <algorithm>
#include <vector>
#include <opencv2/opencv.hpp>

namespace Diasss
{
    class Util
    {
    public:
        static float ComputeIntersection(const std::vector<cv::Mat> &geo_img_s, const std::vector<cv::Mat> &geo_img_t);
    };
}
