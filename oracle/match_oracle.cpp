// TEST INFRASTRUCTURE ONLY -- CPU oracle for the diasss matcher and Frame glue ("ORB mode").
//
// From-scratch restatement of
//   /root/reference/src/core/FEAmatcher.cpp   (RobustMatching :13-50, GeoNearNeighSearch :52-321 with
//                                              USE_SIFT = 0 i.e. the ORB branch :141-176, SCC_x :186-248,
//                                              ConsistentCheck :323-405, DescriptorDistance :442-458)
//   /root/reference/src/core/frame.cpp        (GetNormalizeSSS :57-81, GetFilteredMask :83-124,
//                                              GetGeoImg :126-165, DetectFeature mask filter :184-195)
//   /root/reference/src/util/util.cpp         (ComputeIntersection :13-43)
//   /root/reference/src/core/optimizer.cpp    (GetKpsPairs :575-639, USE_ANNO = 0)
// See orb_oracle.cpp for the rules on who may use this file.  Pinning status: PINNED -- every function here is
// held byte for byte to oracle/_ref (the reference's own FEAmatcher.cpp, frame.cpp, util.cpp:13-43 and
// optimizer.cpp:575-639 compiled unmodified by oracle/build_ref.sh) in tests/test_ref_pin.py: CorresID after
// SCC, every scc push, ConsistentCheck branches, corres_kps rows of both frames, the Frame constructor's
// planes and features, ComputeIntersection, DescriptorDistance, GetKpsPairs, test_demo's frame / pair loop.
// cv::RNG is pinned against cv2.randu (tests/test_oracle_primitives.py).
//
// Documented deviations where the reference is undefined:
//   B3  a direction with no tentative match (ID_loc empty) skips the SCC loop and returns all -1 with
//       an empty scc; ConsistentCheck with an empty scc on either side takes the else-branch
//   B5  GetFilteredMask's 12x12 stamp is clipped to the image and skipped when i<r or j<r
//   B6  mask.at<bool>()==1 is a non-zero byte test
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "oracle_capi.h"

namespace {

struct Rng {  // cv::RNG (core/operations.hpp): multiply-with-carry, default state 0xffffffff
    uint64_t state = 0xffffffffULL;
    uint32_t next() {
        state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
        return (uint32_t)state;
    }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (uint32_t)(b - a) + a); }
};

int descriptor_distance(const uint8_t* a, const uint8_t* b) {  // FEAmatcher.cpp:442-458
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t wa, wb;
        std::memcpy(&wa, a + 4 * i, 4);
        std::memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

void min_max(const double* p, size_t n, double& mn, double& mx) {  // cv::minMaxLoc values
    mn = p[0]; mx = p[0];
    for (size_t i = 1; i < n; i++) { if (p[i] < mn) mn = p[i]; if (p[i] > mx) mx = p[i]; }
}

inline double geo_at(const double* plane, int cols, const orc_keypoint& k) {  // .at<double>(pt.y, pt.x): float -> int truncation
    return plane[(size_t)(int)k.y * cols + (int)k.x];
}

// along-track offset of one tentative match (FEAmatcher.cpp:209-212 / :222-227); float arithmetic, widened on store
inline double track_offset(const orc_keypoint& k, const orc_keypoint& kr, bool flipped, int rows_ref) {
    if (flipped) return std::fabs(k.y - ((float)rows_ref - kr.y + 1));
    return std::fabs(k.y - kr.y);
}

struct SearchResult {
    std::vector<int> corres, pre, best, sec, ncand;
    std::vector<std::pair<int, double>> scc;
};

void geo_nn_search(const orc_frame& f, const orc_frame& ref, SearchResult& r) {  // :52-321
    Rng rng;
    const int n = f.n, nref = ref.n;
    r.corres.assign(n, -1); r.best.assign(n, 1000); r.sec.assign(n, 1000); r.ncand.assign(n, 0);
    r.scc.clear();
    std::vector<int> id_loc;
    const int radius = 8;
    double bx_min, bx_max, by_min, by_max;
    min_max(ref.geo_x, (size_t)ref.rows * ref.cols, bx_min, bx_max);            // :71-72
    min_max(ref.geo_y, (size_t)ref.rows * ref.cols, by_min, by_max);
    const bool flipped = (f.img_id % 2 != ref.img_id % 2);
    std::vector<int> candidate;
    for (int i = 0; i < n; i++) {                                               // :79-183
        const double loc_x = geo_at(f.geo_x, f.cols, f.kps[i]);
        const double loc_y = geo_at(f.geo_y, f.cols, f.kps[i]);
        if (loc_x < bx_min || loc_y < by_min || loc_x > bx_max || loc_y > by_max) continue;
        candidate.clear();
        for (int j = 0; j < nref; j++) {
            const double rx = geo_at(ref.geo_x, ref.cols, ref.kps[j]);
            const double ry = geo_at(ref.geo_y, ref.cols, ref.kps[j]);
            const double geo_dist = std::sqrt((loc_x - rx) * (loc_x - rx) + (loc_y - ry) * (loc_y - ry));
            if (geo_dist < radius) candidate.push_back(j);
        }
        r.ncand[i] = (int)candidate.size();
        if (candidate.empty()) continue;
        int best_dist = 1000, sec_best_dist = 1000, dist_bound = 88;            // :143-147
        if (flipped) dist_bound = 80;
        int best_id = -1;
        const double ratio_test = 0.35;
        for (size_t c = 0; c < candidate.size(); c++) {
            const int d = descriptor_distance(f.desc + 32 * (size_t)i, ref.desc + 32 * (size_t)candidate[c]);
            if (d < best_dist) { sec_best_dist = best_dist; best_dist = d; best_id = candidate[c]; }
            else if (d < sec_best_dist) sec_best_dist = d;
        }
        r.best[i] = best_dist; r.sec[i] = sec_best_dist;
        const double fir_sec_ratio = (double)best_dist / sec_best_dist;
        if (best_id != -1 && best_dist <= dist_bound && fir_sec_ratio <= ratio_test && sec_best_dist != 1000) {
            r.corres[i] = best_id; id_loc.push_back(i);
        } else if (candidate.size() == 1 && best_dist <= dist_bound) {
            r.corres[i] = best_id; id_loc.push_back(i);
        }
    }
    r.pre = r.corres;
    if (id_loc.empty()) {  // B3
        r.corres.assign(n, -1);
        return;
    }
    int final_inlier_num = 0;                                                   // :186-248
    const int max_iter = 1000, sam_num = 2;
    const double PixError = 2.5;
    std::vector<int> final_ids(n, -1), iter_ids(n);
    for (int it = 0; it < max_iter; it++) {
        int cur = 0;
        std::fill(iter_ids.begin(), iter_ids.end(), -1);
        int sampled[2];
        for (int s = 0; s < sam_num; s++) sampled[s] = id_loc[rng.uniform(0, (int)id_loc.size())];
        double ModelX = 0;
        for (int s = 0; s < sam_num; s++)
            ModelX = ModelX + track_offset(f.kps[sampled[s]], ref.kps[r.pre[sampled[s]]], flipped, ref.rows);
        ModelX = ModelX / sam_num;
        for (int j = 0; j < n; j++) {
            if (r.pre[j] == -1) continue;
            const double X_tmp = track_offset(f.kps[j], ref.kps[r.pre[j]], flipped, ref.rows);
            if (std::fabs(ModelX - X_tmp) <= PixError) { iter_ids[j] = r.pre[j]; cur++; }
        }
        if (final_inlier_num < cur) {
            final_ids = iter_ids;
            final_inlier_num = cur;
            r.scc.push_back(std::make_pair(cur, ModelX));
        }
    }
    r.corres = final_ids;
}

// ConsistentCheck (:323-405) on indices: emits (source index, target index) pairs in reference order.
void consistent_check(const orc_frame& s, const orc_frame& t, const std::vector<int>& c1, const std::vector<int>& c2,
                      std::vector<std::pair<int, double>> scc1, std::vector<std::pair<int, double>> scc2,
                      std::vector<std::pair<int, int>>& out) {
    const double kp_diff_thres = 2.5;
    std::sort(scc1.rbegin(), scc1.rend());
    std::sort(scc2.rbegin(), scc2.rend());
    double img_diff = 0;
    if (s.img_id % 2 != t.img_id % 2) img_diff = std::abs(s.rows - t.rows);
    bool merge = false;
    if (!scc1.empty() && !scc2.empty()) {  // B3
        const double kp_diff = std::fabs(std::fabs(scc1[0].second - scc2[0].second) - img_diff);
        merge = kp_diff <= kp_diff_thres;
    }
    if (merge) {
        for (size_t i = 0; i < c1.size(); i++) {
            if (c1[i] == -1) continue;
            if (c2[c1[i]] == (int)i) continue;
            out.push_back({(int)i, c1[i]});
        }
        for (size_t i = 0; i < c2.size(); i++) {
            if (c2[i] == -1) continue;
            out.push_back({c2[i], (int)i});
        }
    } else {
        const int inl1 = (int)(c1.size() - std::count(c1.begin(), c1.end(), -1));
        const int inl2 = (int)(c2.size() - std::count(c2.begin(), c2.end(), -1));
        if (inl1 > inl2) {
            for (size_t i = 0; i < c1.size(); i++) if (c1[i] != -1) out.push_back({(int)i, c1[i]});
        } else {
            for (size_t i = 0; i < c2.size(); i++) if (c2[i] != -1) out.push_back({c2[i], (int)i});
        }
    }
}

}  // namespace

extern "C" {

void orc_rng_draws(uint32_t* out, int n) {
    Rng r;
    for (int i = 0; i < n; i++) out[i] = r.next();
}

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

void orc_geo_nn_search(const orc_frame* f, const orc_frame* ref, int* corres_id, int* pre_corres, int* best_dist,
                       int* sec_dist, int* n_cand, int* scc_count, double* scc_model, int scc_cap, int* n_scc) {
    SearchResult r;
    geo_nn_search(*f, *ref, r);
    const size_t nb = sizeof(int) * (size_t)f->n;
    if (corres_id) std::memcpy(corres_id, r.corres.data(), nb);
    if (pre_corres) std::memcpy(pre_corres, r.pre.data(), nb);
    if (best_dist) std::memcpy(best_dist, r.best.data(), nb);
    if (sec_dist) std::memcpy(sec_dist, r.sec.data(), nb);
    if (n_cand) std::memcpy(n_cand, r.ncand.data(), nb);
    if (n_scc) *n_scc = (int)r.scc.size();
    for (size_t i = 0; i < r.scc.size() && (int)i < scc_cap; i++) {
        if (scc_count) scc_count[i] = r.scc[i].first;
        if (scc_model) scc_model[i] = r.scc[i].second;
    }
}

int orc_robust_matching(const orc_frame* s, const orc_frame* t, double* rows6, int* src_idx, int* tgt_idx, int cap,
                        int* corres1, int* corres2) {  // FEAmatcher.cpp:13-50
    SearchResult r1, r2;
    geo_nn_search(*s, *t, r1);
    geo_nn_search(*t, *s, r2);
    std::vector<std::pair<int, int>> m;
    consistent_check(*s, *t, r1.corres, r2.corres, r1.scc, r2.scc, m);
    if (corres1) std::memcpy(corres1, r1.corres.data(), sizeof(int) * (size_t)s->n);
    if (corres2) std::memcpy(corres2, r2.corres.data(), sizeof(int) * (size_t)t->n);
    const int K = (int)m.size();
    for (int i = 0; i < K && i < cap; i++) {
        const orc_keypoint& a = s->kps[m[i].first];
        const orc_keypoint& b = t->kps[m[i].second];
        if (rows6) {
            double* row = rows6 + 6 * (size_t)i;
            row[0] = s->img_id; row[1] = t->img_id; row[2] = a.y; row[3] = a.x; row[4] = b.y; row[5] = b.x;
        }
        if (src_idx) src_idx[i] = m[i].first;
        if (tgt_idx) tgt_idx[i] = m[i].second;
    }
    return K;
}

// ---- Frame glue -------------------------------------------------------------------------------
int orc_mask_filter(const orc_keypoint* kps, const uint8_t* desc, int n, const uint8_t* mask, int mstep,
                    orc_keypoint* out_kps, uint8_t* out_desc, int* out_index) {  // frame.cpp:184-195 (+B6)
    int m = 0;
    for (int i = 0; i < n; i++) {
        const int v = (int)kps[i].y, u = (int)kps[i].x;
        if (mask[(size_t)v * mstep + u] != 0) {
            if (out_kps) out_kps[m] = kps[i];
            if (out_desc) std::memcpy(out_desc + 32 * (size_t)m, desc + 32 * (size_t)i, 32);
            if (out_index) out_index[m] = i;
            m++;
        }
    }
    return m;
}

void orc_geo_img(int rows, int cols, const double* pose6, const double* g_range, int n_range, double* geo_x,
                 double* geo_y) {  // frame.cpp:126-165 with tf_stb = tf_port = {0,0,0} (:38-39); B4: n_range >= cols/2+1
    (void)n_range;
    const double PI = 3.14159265359;  // frame.cpp:16
    for (int i = 0; i < rows; i++) {
        const double* p = pose6 + 6 * (size_t)i;
        int count = 0;
        for (int j = cols / 2; j < cols; j++) {
            geo_x[(size_t)i * cols + j] = p[3] - 0.0 + g_range[count] * std::cos(p[2] + PI / 2);
            geo_y[(size_t)i * cols + j] = p[4] - 0.0 + g_range[count] * std::sin(p[2] + PI / 2);
            count++;
        }
        for (int j = 0; j < cols / 2; j++) {
            geo_x[(size_t)i * cols + j] = p[3] - 0.0 + g_range[count] * std::cos(p[2] - PI / 2);
            geo_y[(size_t)i * cols + j] = p[4] - 0.0 + g_range[count] * std::sin(p[2] - PI / 2);
            count--;
        }
    }
}

// cv::mean.  OpenCV does not define the summation order (it depends on the SIMD dispatch of the build), so two
// definitions exist here: order 0 = plain sequential sum; order 1 = the two-level 32-lane order the CUDA library
// defines (diasss_b200/csrc/frameprep.cu): per row, lane l adds raw(i,l), raw(i,l+32), ... then the 32 partial sums are
// combined by the butterfly p[l] += p[l^16], ^8, ^4, ^2, ^1; the row sums are combined the same way over rows.
static double butterfly32(double* p) {
    for (int o = 16; o > 0; o >>= 1) {
        double q[32];
        for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ o];
        std::memcpy(p, q, sizeof(q));
    }
    return p[0];
}

double orc_mean(const double* raw, int rows, int cols, int order) {
    const size_t n = (size_t)rows * cols;
    if (order == 0) {
        double sum = 0;
        for (size_t i = 0; i < n; i++) sum += raw[i];
        return sum / (double)n;
    }
    double acc[32];
    for (int l = 0; l < 32; l++) acc[l] = 0.0;
    for (int i = 0; i < rows; i++) {
        double p[32];
        for (int l = 0; l < 32; l++) p[l] = 0.0;
        for (int j = 0; j < cols; j++) p[j & 31] += raw[(size_t)i * cols + j];
        acc[i & 31] += butterfly32(p);
    }
    return butterfly32(acc) / (double)n;
}

void orc_normalize_sss_m(const double* raw, int rows, int cols, double mean, uint8_t* out) {  // frame.cpp:57-81
    const size_t n = (size_t)rows * cols;
    double mn, mx;
    const double max_used = mean * 2.5;
    min_max(raw, n, mn, mx);
    for (size_t i = 0; i < n; i++) {
        double v = (raw[i] - mn) / (max_used - mn) * 255.0;
        if (v > 255.0) v = 255.0;
        int iv = (int)lrint(v);  // convertTo(CV_8U) = saturate_cast<uchar>(cvRound(v))
        out[i] = (uint8_t)std::min(std::max(iv, 0), 255);
    }
}

void orc_filtered_mask_m(const double* raw, int rows, int cols, double mean, uint8_t* out) {  // frame.cpp:83-124 (+B5)
    const float factor = 2.5f;
    const int width = 10, r = 6, side = 150;
    const size_t n = (size_t)rows * cols;
    std::memset(out, 255, n);
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            if (raw[(size_t)i * cols + j] > mean * factor && i >= r && j >= r)
                for (int x = i - r; x < i + r && x < rows; x++)
                    for (int y = j - r; y < j + r && y < cols; y++) out[(size_t)x * cols + y] = 0;
            if (j > cols / 2 - width && j < cols / 2 + width) out[(size_t)i * cols + j] = 0;
            if (i < side || i > rows - side) out[(size_t)i * cols + j] = 0;
            if (j < side * 0.6 || j > cols - side * 0.6) out[(size_t)i * cols + j] = 0;
        }
}

void orc_normalize_sss(const double* raw, int rows, int cols, uint8_t* out) {
    orc_normalize_sss_m(raw, rows, cols, orc_mean(raw, rows, cols, 0), out);
}

void orc_filtered_mask(const double* raw, int rows, int cols, uint8_t* out) {
    orc_filtered_mask_m(raw, rows, cols, orc_mean(raw, rows, cols, 0), out);
}

// Optimizer::GetKpsPairs, USE_ANNO = 0 (src/core/optimizer.cpp:575-639): the rows RobustMatching appended to
// Frame::corres_kps -> [y_s, x_s, slant_s, y_t, x_t, slant_t, 0] per kept correspondence.
int orc_get_kps_pairs(const double* rows6, int k, int id_t, const double* alt_s, const double* gra_s, int n_gra_s, const double* alt_t,
                      const double* gra_t, int n_gra_t, double* out7) {
    int n = 0;
    for (int i = 0; i < k; i++) {
        const double* r = rows6 + (size_t)i * 6;
        const int id_check = (int)r[1];                                                   // :596
        const int ys = (int)r[2], xs = (int)r[3], yt = (int)r[4], xt = (int)r[5];         // :597-598
        const int ds = xs - n_gra_s, dt = xt - n_gra_t;                                   // :603-604
        if (std::abs(ds) < 20 || std::abs(dt) < 20) continue;                             // :605-609 nadir
        if (id_check != id_t) continue;                                                   // :613
        double* q = out7 + (size_t)n * 7;
        q[0] = ys; q[1] = xs;
        q[2] = std::sqrt(alt_s[ys] * alt_s[ys] + gra_s[std::abs(ds)] * gra_s[std::abs(ds)]);   // :617
        q[3] = yt; q[4] = xt;
        q[5] = std::sqrt(alt_t[yt] * alt_t[yt] + gra_t[std::abs(dt)] * gra_t[std::abs(dt)]);   // :619
        q[6] = 0.0;
        n++;
    }
    return n;
}

float orc_compute_intersection(const double* sx, const double* sy, int sn, const double* tx, const double* ty,
                               int tn) {  // util.cpp:13-43 (areas and ratio evaluated in float)
    float output = 0.0f;
    double sx_min, sx_max, sy_min, sy_max, tx_min, tx_max, ty_min, ty_max;
    min_max(sx, sn, sx_min, sx_max); min_max(sy, sn, sy_min, sy_max);
    min_max(tx, tn, tx_min, tx_max); min_max(ty, tn, ty_min, ty_max);
    float x_dist_ol = (float)(std::min(sx_max, tx_max) - std::max(sx_min, tx_min));
    float y_dist_ol = (float)(std::min(ty_max, sy_max) - std::max(sy_min, ty_min));
    if (x_dist_ol > 0 && y_dist_ol > 0) {
        float area_ol = x_dist_ol * y_dist_ol;
        float area_s = (float)(std::abs(sx_max - sx_min) * std::abs(sy_max - sy_min));
        float area_t = (float)(std::abs(tx_max - tx_min) * std::abs(ty_max - ty_min));
        output = area_ol / (area_s + area_t - area_ol);
    }
    return output;
}

}  // extern "C"
