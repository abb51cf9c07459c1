// C++ host-side check of include/diasss_b200/shim.hpp: uses ORB_SLAM2::ORBextractor and Diasss::FEAmatcher exactly
// the way the reference does (frame.cpp:180-195, diasss2.cpp:88-97, FEAmatcher.cpp:13-50) on frames dumped by
// tests/test_cpp_shim.py, and writes every result to a binary file that the Python test compares with the oracle.
//
//   test_shim <in.bin> <out.bin>
//
// in.bin : int32 n_frames; per frame: int32 img_id, rows, cols; u8 image[rows*cols]; u8 mask[rows*cols];
//          f64 geo_x[rows*cols]; f64 geo_y[rows*cols]
// out.bin: per frame: int32 n_all (operator() output), int32 n (after the mask filter), KeyPoint kps[n], u8 desc[n*32];
//          then for the pair (0,1): int32 K, f64 source corres_kps[K*6], f64 target corres_kps[K*6];
//          int32 n0, int32 CorresID_1[n0], int32 n1, int32 CorresID_2[n1], int32 K2, (KeyPoint,KeyPoint)[K2] from
//          ConsistentCheck, int32 DescriptorDistance(dst0.row(0), dst1.row(0)); then the same via DetectFeatureB200: int32 n, kps.
#define DSX_SHIM_MINI_CV "../../tests/cpp/mini_cv.hpp"
#include "../../include/diasss_b200/shim.hpp"

#include <cstdio>
#include <fstream>
#include <iostream>

namespace Diasss {
// the fields of Diasss::Frame the path touches (src/core/frame.h:30-46)
struct Frame {
    int img_id;
    cv::Mat norm_img, flt_mask;
    std::vector<cv::Mat> geo_img;
    std::vector<cv::KeyPoint> kps;
    cv::Mat dst;
    cv::Mat corres_kps;

    // Frame::DetectFeature, statement for statement (frame.cpp:167-203)
    void DetectFeature(const cv::Mat& img, const cv::Mat& mask, std::vector<cv::KeyPoint>& kps_, cv::Mat& dst_, int* n_all) {
        std::vector<cv::KeyPoint> keypoints;
        cv::Mat descriptors;
        ORB_SLAM2::ORBextractor orb = ORB_SLAM2::ORBextractor(2000, 1.2, 6, 12, 7);
        orb(img, cv::Mat(), keypoints, descriptors);
        *n_all = (int)keypoints.size();
        for (size_t i = 0; i < keypoints.size(); i++) {
            const int v = keypoints[i].pt.y;
            const int u = keypoints[i].pt.x;
            if (mask.at<unsigned char>(v, u) != 0) {   // B6
                kps_.push_back(keypoints[i]);
                dst_.push_back(descriptors.row(i));
            }
        }
    }
};
}  // namespace Diasss

template <typename T> static void rd(std::ifstream& f, T* p, size_t n) { f.read(reinterpret_cast<char*>(p), sizeof(T) * n); }
template <typename T> static void wr(std::ofstream& f, const T* p, size_t n) { f.write(reinterpret_cast<const char*>(p), sizeof(T) * n); }

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: test_shim in.bin out.bin\n"); return 2; }
    try {
        std::ifstream in(argv[1], std::ios::binary);
        if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
        std::ofstream out(argv[2], std::ios::binary);
        int32_t nf = 0;
        rd(in, &nf, 1);
        std::vector<Diasss::Frame> frames((size_t)nf);
        for (int i = 0; i < nf; i++) {
            Diasss::Frame& F = frames[i];
            int32_t h[3];
            rd(in, h, 3);
            F.img_id = h[0];
            F.norm_img.create(h[1], h[2], CV_8U); F.flt_mask.create(h[1], h[2], CV_8U);
            rd(in, F.norm_img.data, (size_t)h[1] * h[2]); rd(in, F.flt_mask.data, (size_t)h[1] * h[2]);
            F.geo_img.resize(2);
            for (int q = 0; q < 2; q++) { F.geo_img[q].create(h[1], h[2], CV_64F); rd(in, F.geo_img[q].ptr<double>(), (size_t)h[1] * h[2]); }
            int32_t n_all = 0;
            F.DetectFeature(F.norm_img, F.flt_mask, F.kps, F.dst, &n_all);
            const int32_t n = (int32_t)F.kps.size();
            wr(out, &n_all, 1); wr(out, &n, 1);
            wr(out, F.kps.data(), (size_t)n);
            for (int r = 0; r < n; r++) wr(out, F.dst.ptr<uint8_t>(r), 32);
        }
        if (nf >= 2) {
            Diasss::Frame &A = frames[0], &B = frames[1];
            Diasss::FEAmatcher::RobustMatching(A, B);                             // diasss2.cpp:95
            const int32_t K = A.corres_kps.rows;
            wr(out, &K, 1);
            for (int r = 0; r < K; r++) wr(out, A.corres_kps.ptr<double>(r), 6);
            for (int r = 0; r < K; r++) wr(out, B.corres_kps.ptr<double>(r), 6);
            // the pieces RobustMatching is made of, called the way FEAmatcher.cpp:26-32 calls them
            std::vector<std::pair<int, double>> scc_1, scc_2;
            std::vector<int> c1 = Diasss::FEAmatcher::GeoNearNeighSearch(A.img_id, B.img_id, A.norm_img, B.norm_img, A.kps, A.dst, A.geo_img,
                                                                         B.kps, B.dst, B.geo_img, scc_1);
            std::vector<int> c2 = Diasss::FEAmatcher::GeoNearNeighSearch(B.img_id, A.img_id, B.norm_img, A.norm_img, B.kps, B.dst, B.geo_img,
                                                                         A.kps, A.dst, A.geo_img, scc_2);
            std::vector<cv::KeyPoint> SourceKeys, TargetKeys;
            Diasss::FEAmatcher::ConsistentCheck(A, B, c1, c2, scc_1, scc_2, SourceKeys, TargetKeys);
            int32_t n0 = (int32_t)c1.size(), n1 = (int32_t)c2.size(), K2 = (int32_t)SourceKeys.size();
            wr(out, &n0, 1); wr(out, c1.data(), c1.size());
            wr(out, &n1, 1); wr(out, c2.data(), c2.size());
            wr(out, &K2, 1);
            for (int i = 0; i < K2; i++) { wr(out, &SourceKeys[i], 1); wr(out, &TargetKeys[i], 1); }
            int32_t dd = (A.dst.rows && B.dst.rows) ? Diasss::FEAmatcher::DescriptorDistance(A.dst.row(0), B.dst.row(0)) : -1;
            wr(out, &dd, 1);
            // the one-call DetectFeature replacement
            std::vector<cv::KeyPoint> k2; cv::Mat d2;
            Diasss::DetectFeatureB200(A.norm_img, A.flt_mask, k2, d2);
            int32_t n2 = (int32_t)k2.size();
            wr(out, &n2, 1); wr(out, k2.data(), k2.size());
            for (int r = 0; r < n2; r++) wr(out, d2.ptr<uint8_t>(r), 32);
        }
        // empty image: silent return, no keypoints (ORBextractor.cpp:1052)
        {
            ORB_SLAM2::ORBextractor orb(2000, 1.2f, 6, 12, 7);
            std::vector<cv::KeyPoint> k; cv::Mat d;
            orb(cv::Mat(), cv::Mat(), k, d);
            int32_t e = (int32_t)k.size() + (d.empty() ? 0 : 1000);
            wr(out, &e, 1);
            int32_t lv = orb.GetLevels(); float sf = orb.GetScaleFactor();
            wr(out, &lv, 1); wr(out, &sf, 1);
            std::vector<float> s = orb.GetScaleFactors(), is = orb.GetInverseScaleFactors(), s2 = orb.GetScaleSigmaSquares(),
                               is2 = orb.GetInverseScaleSigmaSquares();
            wr(out, s.data(), s.size()); wr(out, is.data(), is.size()); wr(out, s2.data(), s2.size()); wr(out, is2.data(), is2.size());
        }
        out.close();
        std::printf("test_shim ok\n");
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "test_shim failed: %s\n", e.what());
        return 3;
    }
}
