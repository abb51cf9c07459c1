// diasss_b200/shim.hpp -- header-only C++ drop-in for the two classes of halajun/diasss that sit on the hot path.
//
//   ORB_SLAM2::ORBextractor      replaces thirdparty/ORBextractor.{h,cpp}  (ctor :410-470, operator() :1049-1113)
//   Diasss::FEAmatcher           replaces src/core/FEAmatcher.{h,cpp}      (RobustMatching :13-50,
//                                GeoNearNeighSearch :52-321, ConsistentCheck :323-405, DescriptorDistance :442-458)
//
// Same class names, member names, argument order and meaning as the reference, so that src/core/frame.cpp
// (Frame::DetectFeature, :180-181) and src/diasss2.cpp (the i<j RobustMatching loop, :88-97) compile against
// it unchanged; see INTEGRATION.md.  Every numeric step is a call into libdiasss_b200.so (CUDA, sm_100a) through
// the C ABI of include/diasss_b200.h -- this header only converts between cv:: containers and plain buffers.
// Semantics are the reference's ORB mode (rBRIEF + Hamming) with the Appendix-B definitions.
//
// Error behaviour: an empty image returns silently with no keypoints (ORBextractor.cpp:1052);
// a non-CV_8UC1 image trips an assert (:1056); any C-ABI failure -- including "no CUDA device": there is no CPU
// fallback -- throws std::runtime_error carrying dsx_last_error().
//
// Build against OpenCV as usual; define DSX_SHIM_MINI_CV before including to use a stand-in for the few cv types
// (tests/cpp/mini_cv.hpp, only for images without OpenCV's C++ headers).
#pragma once

#ifdef DSX_SHIM_MINI_CV
#include DSX_SHIM_MINI_CV
#else
#include <opencv2/core.hpp>
#endif

#include <algorithm>
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../diasss_b200.h"

namespace dsx_shim {

inline void check(int status, const char* what) {
    if (status != DSX_OK)
        throw std::runtime_error(std::string(what) + ": diasss_b200 status " + std::to_string(status) + ": " + dsx_last_error());
}

// One context per parameter set and thread; the reference constructs an extractor per frame (frame.cpp:180),
// so construction must be cheap: contexts are cached.
struct ContextKey {
    int nfeatures, nlevels, ini, min;
    float scale;
    bool operator==(const ContextKey& o) const {
        return nfeatures == o.nfeatures && nlevels == o.nlevels && ini == o.ini && min == o.min && scale == o.scale;
    }
};

class ContextCache {
public:
    static ContextCache& instance() {
        static thread_local ContextCache cache;
        return cache;
    }
    static dsx_ctx* get(const ContextKey& k) {
        ContextCache& cache = instance();
        for (auto& e : cache.entries_)
            if (e.first == k) return e.second;
        dsx_params p;
        dsx_default_params(&p);
        p.nfeatures = k.nfeatures; p.scale_factor = k.scale; p.nlevels = k.nlevels; p.ini_th_fast = k.ini; p.min_th_fast = k.min;
        dsx_ctx* ctx = nullptr;
        check(dsx_create(&p, nullptr, &ctx), "dsx_create");
        cache.entries_.push_back(std::make_pair(k, ctx));
        return ctx;
    }
    // The matcher's literals do not depend on the extractor's parameters, but a context's host-match staging is sized by
    // its extractor capacity (dsx_max_keypoints): pick the roomiest cached context and, if a frame carries more
    // keypoints than that (an extractor built with a larger nfeatures elsewhere), create one that holds `need`.
    // Default parameters = frame.cpp:180.
    static dsx_ctx* matcher(int need = 0) {
        ContextCache& cache = instance();
        dsx_ctx* best = nullptr;
        for (auto& e : cache.entries_)
            if (!best || dsx_max_keypoints(e.second) > dsx_max_keypoints(best)) best = e.second;
        if (!best) best = get(ContextKey{2000, 6, 12, 7, 1.2f});
        if (dsx_max_keypoints(best) < need) best = get(ContextKey{need, 6, 12, 7, 1.2f});
        return best;
    }
    ~ContextCache() { for (auto& e : entries_) dsx_destroy(e.second); }
private:
    std::vector<std::pair<ContextKey, dsx_ctx*>> entries_;
};

static_assert(sizeof(cv::KeyPoint) == sizeof(dsx_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI uses");

// Frame fields -> dsx_frame (FEAmatcher.cpp:81-85 look-ups, :71-72 bbox); `store` keeps the converted buffers alive
struct FrameBuffers {
    std::vector<double> geo_xy;
    cv::Mat desc;   // continuous copy if needed
    dsx_frame f;
};

inline void fill_frame(int img_id, int rows, const std::vector<cv::KeyPoint>& kps, const cv::Mat& dst,
                       const std::vector<cv::Mat>& geo_img, FrameBuffers& B) {
    const int n = (int)kps.size();
    B.geo_xy.assign((size_t)2 * (n > 0 ? n : 1), 0.0);
    B.desc = dst.isContinuous() ? dst : dst.clone();
    B.f.img_id = img_id; B.f.rows = rows; B.f.n = n;
    B.f.kps = reinterpret_cast<const dsx_keypoint*>(kps.data());
    B.f.desc = B.desc.data;
    B.f.geo_xy = B.geo_xy.data();
    assert(geo_img.size() >= 2 && geo_img[0].type() == CV_64F);
    check(dsx_frame_geo_from_planes(B.f.kps, n, geo_img[0].ptr<double>(), geo_img[1].ptr<double>(), geo_img[0].rows,
                                    geo_img[0].cols, geo_img[0].step / sizeof(double), B.geo_xy.data(), B.f.bbox),
          "dsx_frame_geo_from_planes");
}

}  // namespace dsx_shim

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_) {
        ctx_ = dsx_shim::ContextCache::get(dsx_shim::ContextKey{nfeatures_, nlevels_, iniThFAST_, minThFAST_, scaleFactor_});
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
        mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels); umax.resize(16);
        dsx_shim::check(dsx_get_tables(ctx_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                       mvInvLevelSigma2.data(), mnFeaturesPerLevel.data(), umax.data()), "dsx_get_tables");
    }
    ~ORBextractor() {}

    // Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference (ORBextractor.h:58).
    void operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                    cv::OutputArray _descriptors) {
        if (_image.empty()) return;                                              // ORBextractor.cpp:1052-1053
        cv::Mat image = _image.getMat();
        assert(image.type() == CV_8UC1);                                         // :1056
        const int cap = dsx_max_keypoints(ctx_);
        std::vector<cv::KeyPoint> kps((size_t)cap);
        std::vector<uint8_t> desc((size_t)cap * DSX_DESC_BYTES);
        int n = 0;
        dsx_shim::check(dsx_extract(ctx_, image.data, image.rows, image.cols, image.step,
                                    reinterpret_cast<dsx_keypoint*>(kps.data()), desc.data(), cap, &n), "dsx_extract");
        if (n == 0) _descriptors.release();                                      // :1070-1071
        else {
            _descriptors.create(n, 32, CV_8U);                                   // :1074
            cv::Mat d = _descriptors.getMat();
            for (int i = 0; i < n; i++) std::memcpy(d.ptr<uint8_t>(i), desc.data() + (size_t)i * 32, 32);
        }
        _keypoints.clear();                                                      // :1078
        _keypoints.assign(kps.begin(), kps.begin() + n);
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Public in the reference (ORBextractor.h:85) but read by nothing outside the extractor; the pyramid lives in
    // device memory here, so this stays empty.
    std::vector<cv::Mat> mvImagePyramid;

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;
    dsx_ctx* ctx_;
};

}  // namespace ORB_SLAM2

namespace Diasss {

// Frame::DetectFeature's body (frame.cpp:167-203) as one call: operator() + the mask filter on the device.
// A maintainer can replace the body of Frame::DetectFeature by `Diasss::DetectFeatureB200(img, mask, kps, dst);`
// (optional: leaving frame.cpp untouched already routes operator() through the GPU).
inline void DetectFeatureB200(const cv::Mat& img, const cv::Mat& mask, std::vector<cv::KeyPoint>& kps, cv::Mat& dst) {
    if (img.empty()) return;
    dsx_ctx* ctx = dsx_shim::ContextCache::get(dsx_shim::ContextKey{2000, 6, 12, 7, 1.2f});   // frame.cpp:180
    const int cap = dsx_max_keypoints(ctx);
    std::vector<cv::KeyPoint> k((size_t)cap);
    std::vector<uint8_t> d((size_t)cap * DSX_DESC_BYTES);
    int n = 0;
    dsx_shim::check(dsx_detect_feature(ctx, img.data, img.step, mask.data, mask.step, img.rows, img.cols,
                                       reinterpret_cast<dsx_keypoint*>(k.data()), d.data(), cap, &n), "dsx_detect_feature");
    for (int i = 0; i < n; i++) {                                                // frame.cpp:190-191
        kps.push_back(k[i]);
        dst.push_back(cv::Mat(1, 32, CV_8U, d.data() + (size_t)i * 32));
    }
}

class FEAmatcher {
public:
    // FrameT = Diasss::Frame (src/core/frame.h:30-46); a template only so that this header does not need frame.h.
    template <class FrameT>
    static void RobustMatching(FrameT& SourceFrame, FrameT& TargetFrame) {
        dsx_ctx* ctx = dsx_shim::ContextCache::matcher((int)std::max(SourceFrame.kps.size(), TargetFrame.kps.size()));
        dsx_shim::FrameBuffers S, T;
        dsx_shim::fill_frame(SourceFrame.img_id, SourceFrame.norm_img.rows, SourceFrame.kps, SourceFrame.dst, SourceFrame.geo_img, S);
        dsx_shim::fill_frame(TargetFrame.img_id, TargetFrame.norm_img.rows, TargetFrame.kps, TargetFrame.dst, TargetFrame.geo_img, T);
        const int cap = S.f.n + T.f.n;
        std::vector<double> rows6((size_t)6 * (cap > 0 ? cap : 1));
        int k = 0;
        dsx_shim::check(dsx_robust_matching(ctx, &S.f, &T.f, rows6.data(), nullptr, nullptr, cap, &k), "dsx_robust_matching");
        for (int i = 0; i < k; i++) {                                            // FEAmatcher.cpp:35-45
            const double* r = rows6.data() + (size_t)6 * i;
            cv::Mat kp_pair_s(1, 6, CV_64F), kp_pair_t(1, 6, CV_64F);
            const double s6[6] = {r[0], r[1], r[2], r[3], r[4], r[5]}, t6[6] = {r[1], r[0], r[4], r[5], r[2], r[3]};
            for (int c = 0; c < 6; c++) { kp_pair_s.template at<double>(0, c) = s6[c]; kp_pair_t.template at<double>(0, c) = t6[c]; }
            SourceFrame.corres_kps.push_back(kp_pair_s);
            TargetFrame.corres_kps.push_back(kp_pair_t);
        }
    }

    static std::vector<int> GeoNearNeighSearch(const int& img_id, const int& img_id_ref, const cv::Mat& img, const cv::Mat& img_ref,
                                               const std::vector<cv::KeyPoint>& kps, const cv::Mat& dst,
                                               const std::vector<cv::Mat>& geo_img, const std::vector<cv::KeyPoint>& kps_ref,
                                               const cv::Mat& dst_ref, const std::vector<cv::Mat>& geo_img_ref,
                                               std::vector<std::pair<int, double>>& scc) {
        dsx_ctx* ctx = dsx_shim::ContextCache::matcher((int)std::max(kps.size(), kps_ref.size()));
        dsx_shim::FrameBuffers F, R;
        dsx_shim::fill_frame(img_id, img.rows, kps, dst, geo_img, F);
        dsx_shim::fill_frame(img_id_ref, img_ref.rows, kps_ref, dst_ref, geo_img_ref, R);
        std::vector<int> CorresID(kps.size(), -1);
        int32_t cnt = 0; double model = 0;
        dsx_shim::check(dsx_geo_near_neigh_search(ctx, &F.f, &R.f, CorresID.empty() ? nullptr : CorresID.data(), &cnt, &model),
                        "dsx_geo_near_neigh_search");
        // the reference appends every strictly-better (count, ModelX) (:237-242); only its maximum is read afterwards
        // (ConsistentCheck sorts descending and takes [0], :331-344), so the maximum alone is returned.
        if (cnt > 0) scc.push_back(std::make_pair((int)cnt, model));
        return CorresID;
    }

    template <class FrameT>
    static void ConsistentCheck(const FrameT& SourceFrame, const FrameT& TargetFrame, const std::vector<int>& CorresID_1,
                                const std::vector<int>& CorresID_2, std::vector<std::pair<int, double>>& scc_1,
                                std::vector<std::pair<int, double>>& scc_2, std::vector<cv::KeyPoint>& SourceKeys,
                                std::vector<cv::KeyPoint>& TargetKeys) {
        dsx_ctx* ctx = dsx_shim::ContextCache::matcher((int)std::max(CorresID_1.size(), CorresID_2.size()));
        auto best = [](const std::vector<std::pair<int, double>>& s, int32_t& c, double& m) {
            c = 0; m = 0;                                                        // max (count, ModelX) == sorted[0] (:331-332)
            for (auto& e : s) if (c == 0 || e.first > c || (e.first == c && e.second > m)) { c = e.first; m = e.second; }
        };
        int32_t c1, c2; double m1, m2;
        best(scc_1, c1, m1); best(scc_2, c2, m2);
        const int cap = (int)(CorresID_1.size() + CorresID_2.size());
        std::vector<int32_t> si((size_t)(cap > 0 ? cap : 1)), ti((size_t)(cap > 0 ? cap : 1));
        int k = 0;
        dsx_shim::check(dsx_consistent_check(ctx, SourceFrame.img_id, SourceFrame.norm_img.rows, (int)CorresID_1.size(),
                                             TargetFrame.img_id, TargetFrame.norm_img.rows, (int)CorresID_2.size(),
                                             CorresID_1.data(), CorresID_2.data(), c1, m1, c2, m2, si.data(), ti.data(), cap, &k),
                        "dsx_consistent_check");
        for (int i = 0; i < k; i++) {
            SourceKeys.push_back(SourceFrame.kps[si[i]]);
            TargetKeys.push_back(TargetFrame.kps[ti[i]]);
        }
    }

    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        int32_t d = 0;
        dsx_shim::check(dsx_descriptor_distance(dsx_shim::ContextCache::matcher(), a.data, b.data, 1, &d), "dsx_descriptor_distance");
        return d;
    }
};

}  // namespace Diasss
