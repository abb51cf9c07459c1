// TEST INFRASTRUCTURE ONLY -- C interface around the REFERENCE'S OWN hot-path code, compiled unmodified from
// /root/reference (ORBextractor.cpp, FEAmatcher.cpp, frame.cpp, util.cpp:1-43) against the OpenCV stand-in of this
// directory.  Built into oracle/_ref/ by oracle/build_ref.sh; loaded through ctypes by oracle/ref.py.  It is the
// pin for the oracle's control flow (cell loop, quadtree, matcher, SCC, merge) and the "reference" CPU arm of
// bench.py.  Nothing in the product links or loads it.
//
// This file adds no algorithm: it constructs the reference's objects, calls their public (or, through a deriving
// probe class, protected) members and copies the results out.  The only loop restated here is test_demo's
// frame / pair loop (src/diasss2.cpp:82-97; diasss2.cpp itself needs GTSAM to compile), spread over host threads.
//
// Two run-time knobs select how the reference's two host-dependent behaviours resolve:
//   ref_set_heap_mode(1)  std::list<ExtractorNode> nodes come from a bump allocator, so that node addresses grow
//                         in creation order (ORBextractor.cpp:684 sorts by address).  0 = the process heap (glibc).
//   ref_set_libm_mode(1)  cosf / sinf (ORBextractor.cpp:113) are evaluated in double and rounded (the oracle's
//                         definition A5).  0 = the platform's libm.
#include <atomic>
#include <cstdio>
#include <functional>
#include <list>
#include <mutex>
#include <new>
#include <thread>

#include <dlfcn.h>
#include <sys/mman.h>

#include "ORBextractor.h"
#include "frame.h"
#include "FEAmatcher.h"
#include "util.h"
#include "optimizer_getkpspairs.h"
}   // (the stub header leaves namespace Diasss open for the function body build_ref.sh appends to it)

#include "../oracle_capi.h"

#ifndef REF_BUILD_INFO
#define REF_BUILD_INFO "unknown"
#endif

// ---------------------------------------------------------------------------------------------------------------
// heap hook (bound inside this library only: -Wl,-Bsymbolic)
// ---------------------------------------------------------------------------------------------------------------
namespace {
std::atomic<int> g_heap_monotone{0};
std::atomic<int> g_libm_a5{0};
const size_t kNodeBytes = sizeof(std::_List_node<ORB_SLAM2::ExtractorNode>);
const size_t kArenaBytes = (size_t)1 << 36;   // address space only (MAP_NORESERVE); pages are touched as used
char* g_arena = nullptr;
std::atomic<size_t> g_arena_off{0};
std::once_flag g_arena_once;

void* arena_alloc(size_t n) {
    std::call_once(g_arena_once, [] {
        void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::perror("oracle/_ref arena"); std::abort(); }
        g_arena = (char*)p;
    });
    const size_t sz = (n + 15) & ~(size_t)15;
    const size_t off = g_arena_off.fetch_add(sz);
    if (off + sz > kArenaBytes) { std::fprintf(stderr, "oracle/_ref arena exhausted\n"); std::abort(); }
    return g_arena + off;
}
inline bool in_arena(void* p) { return g_arena && (char*)p >= g_arena && (char*)p < g_arena + kArenaBytes; }
inline void* heap_new(size_t n) {
    if (n == kNodeBytes && g_heap_monotone.load(std::memory_order_relaxed)) return arena_alloc(n);
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
inline void heap_delete(void* p) {
    if (!p || in_arena(p)) return;
    std::free(p);
}
}  // namespace

void* operator new(size_t n) { return heap_new(n); }
void* operator new[](size_t n) { return heap_new(n); }
void* operator new(size_t n, const std::nothrow_t&) noexcept { try { return heap_new(n); } catch (...) { return nullptr; } }
void* operator new[](size_t n, const std::nothrow_t&) noexcept { try { return heap_new(n); } catch (...) { return nullptr; } }
void operator delete(void* p) noexcept { heap_delete(p); }
void operator delete[](void* p) noexcept { heap_delete(p); }
void operator delete(void* p, size_t) noexcept { heap_delete(p); }
void operator delete[](void* p, size_t) noexcept { heap_delete(p); }
void operator delete(void* p, const std::nothrow_t&) noexcept { heap_delete(p); }
void operator delete[](void* p, const std::nothrow_t&) noexcept { heap_delete(p); }

// ---------------------------------------------------------------------------------------------------------------
// libm hook: the float cos / sin of computeOrbDescriptor (ORBextractor.cpp:113); gcc may fuse the pair into sincosf
// ---------------------------------------------------------------------------------------------------------------
extern "C" {
float cosf(float x) noexcept {
    if (g_libm_a5.load(std::memory_order_relaxed)) return (float)cos((double)x);
    static float (*real)(float) = (float (*)(float))dlsym(RTLD_NEXT, "cosf");
    return real(x);
}
float sinf(float x) noexcept {
    if (g_libm_a5.load(std::memory_order_relaxed)) return (float)sin((double)x);
    static float (*real)(float) = (float (*)(float))dlsym(RTLD_NEXT, "sinf");
    return real(x);
}
void sincosf(float x, float* s, float* c) noexcept {
    if (g_libm_a5.load(std::memory_order_relaxed)) { *s = (float)sin((double)x); *c = (float)cos((double)x); return; }
    static void (*real)(float, float*, float*) = (void (*)(float, float*, float*))dlsym(RTLD_NEXT, "sincosf");
    real(x, s, c);
}
}

// ---------------------------------------------------------------------------------------------------------------
namespace {

using ORB_SLAM2::ORBextractor;

struct Probe : public ORBextractor {   // reaches the protected members; adds no behaviour
    Probe(int nf, float sf, int nl, int ini, int mn) : ORBextractor(nf, sf, nl, ini, mn) {}
    using ORBextractor::DistributeOctTree;
    using ORBextractor::mnFeaturesPerLevel;
    using ORBextractor::umax;
    using ORBextractor::mvScaleFactor;
    using ORBextractor::mvInvScaleFactor;
    using ORBextractor::nlevels;
    std::vector<std::vector<int>> candidates;   // per level (x, y, response), coordinates relative to (16,16)
    std::vector<cv::KeyPoint> last_kps;
};

void copy_out(const std::vector<cv::KeyPoint>& k, orc_keypoint* out) { std::memcpy(out, k.data(), k.size() * sizeof(cv::KeyPoint)); }

// Turns the cv::FAST call log of one operator() run into the per-level candidate lists the reference builds in
// vToDistributeKeys (ORBextractor.cpp:818-826: a cell contributes the result of its last, non-empty call).
void candidates_from_log(Probe* e, const std::vector<cv::stub::FastCall>& log) {
    e->candidates.assign(e->nlevels, {});
    for (const auto& c : log) {
        if (c.out.empty()) continue;
        for (int l = 0; l < e->nlevels; l++) {
            const cv::Mat& im = e->mvImagePyramid[l];
            if (c.data < im.data || c.data >= im.data + (size_t)im.rows * im.step) continue;
            const size_t off = (size_t)(c.data - im.data);
            const int y0 = (int)(off / im.step), x0 = (int)(off % im.step);
            for (const auto& k : c.out) {
                e->candidates[l].push_back((int)k.pt.x + x0 - 16);
                e->candidates[l].push_back((int)k.pt.y + y0 - 16);
                e->candidates[l].push_back((int)k.response);
            }
            break;
        }
    }
}

cv::Mat wrap_f64(const double* p, int rows, int cols) { return cv::Mat(rows, cols, CV_64F, (void*)p); }

// A Diasss::Frame whose constructor has run on a blank 200x200 swath (Frame has no other constructor,
// frame.h:18-20); the matcher-level entry points below then set the public fields the matcher reads.
const Diasss::Frame& blank_frame() {
    static Diasss::Frame* f = [] {
        const int n = 200;
        cv::Mat img = cv::Mat::zeros(n, n, CV_64F), pose = cv::Mat::zeros(n, 6, CV_64F), anno;
        std::vector<double> alt(n, 1.0), gr(n / 2 + 1, 0.0);
        return new Diasss::Frame(0, img, pose, alt, gr, anno);
    }();
    return *f;
}

Diasss::Frame make_frame(const orc_frame* in) {
    Diasss::Frame f = blank_frame();
    f.img_id = in->img_id;
    f.norm_img = cv::Mat(in->rows, in->cols, CV_8U, (void*)nullptr, (size_t)in->cols);   // only .rows is read
    f.kps.assign((const cv::KeyPoint*)in->kps, (const cv::KeyPoint*)in->kps + in->n);
    f.dst = in->n ? cv::Mat(in->n, 32, CV_8U, (void*)in->desc) : cv::Mat();
    f.geo_img.clear();
    f.geo_img.push_back(wrap_f64(in->geo_x, in->rows, in->cols));
    f.geo_img.push_back(wrap_f64(in->geo_y, in->rows, in->cols));
    f.corres_kps = cv::Mat();
    return f;
}

// What Frame::Frame does after GetNormalizeSSS / GetFilteredMask (frame.cpp:49-52), through the reference's own public
// members, for a frame whose norm_img and flt_mask are given (the synthetic surveys are rendered as u8 images).
void prepare_from_images(Diasss::Frame& f, int id, const cv::Mat& norm, const cv::Mat& mask, const cv::Mat& pose,
                         const std::vector<double>& gr) {
    f.img_id = id;
    f.norm_img = norm;
    f.flt_mask = mask;
    f.dr_poses = pose;
    f.ground_ranges = gr;
    f.geo_img = f.GetGeoImg(norm.rows, norm.cols, pose, gr, f.tf_stb, f.tf_port);
    f.kps.clear();
    f.dst = cv::Mat();
    f.corres_kps = cv::Mat();
    f.DetectFeature(f.norm_img, f.flt_mask, f.kps, f.dst);
}

void parallel_for(int n, int threads, const std::function<void(int)>& fn) {
    if (threads <= 1 || n <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < std::min(threads, n); t++)
        pool.emplace_back([&] { for (int i; (i = next.fetch_add(1)) < n;) fn(i); });
    for (auto& t : pool) t.join();
}

struct Survey {
    int threads = 1;
    struct Input {
        int id, rows, cols;
        const double* raw;            // f64 waterfall (full constructor), or null:
        const uint8_t *norm, *mask;   // prepared norm_img / flt_mask (GetGeoImg + DetectFeature only)
        const double* pose;
        std::vector<double> alt, gr;
    };
    std::vector<Input> inputs;
    std::vector<Diasss::Frame*> frames;
    std::vector<std::pair<int, int>> pairs;     // every i<j in loop order
    std::vector<float> overlap;                 // ComputeIntersection per i<j
    std::vector<int> matched;                   // 1 if RobustMatching ran
    std::vector<cv::Mat> pair_rows;             // K x 6 CV_64F per pair: what the pair appended to Source.corres_kps
    ~Survey() { for (auto* f : frames) delete f; }
};

}  // namespace

extern "C" {

const char* ref_build_info() { return REF_BUILD_INFO; }
void ref_set_heap_mode(int monotone) { g_heap_monotone.store(monotone); }
void ref_set_libm_mode(int a5) { g_libm_a5.store(a5); }
void ref_set_mean_order(int order) { cv::stub::mean_order = order; }
int ref_node_bytes() { return (int)kNodeBytes; }

// ---- ORB_SLAM2::ORBextractor -----------------------------------------------------------------------------------
void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    return new Probe(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_extractor_destroy(void* h) { delete (Probe*)h; }

void ref_extractor_tables(void* h, float* scale, float* inv_scale, int* features_per_level, int* umax16) {
    Probe* e = (Probe*)h;
    std::vector<float> s = e->GetScaleFactors(), is = e->GetInverseScaleFactors();
    for (int i = 0; i < e->GetLevels(); i++) { scale[i] = s[i]; inv_scale[i] = is[i]; features_per_level[i] = e->mnFeaturesPerLevel[i]; }
    for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

// operator()(image, cv::Mat(), keypoints, descriptors), as Frame::DetectFeature calls it (frame.cpp:181)
int ref_extractor_run(void* h, const uint8_t* img, int rows, int cols, int step, orc_keypoint* kps, uint8_t* desc, int cap) {
    Probe* e = (Probe*)h;
    cv::Mat image(rows, cols, CV_8U, (void*)img, (size_t)step), descriptors;
    std::vector<cv::KeyPoint> keypoints;
    std::vector<cv::stub::FastCall> log;
    cv::stub::fast_log = &log;
    (*e)(image, cv::Mat(), keypoints, descriptors);
    cv::stub::fast_log = nullptr;
    candidates_from_log(e, log);
    const int n = (int)keypoints.size();
    if (n <= cap) {
        copy_out(keypoints, kps);
        for (int i = 0; i < n; i++) std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    return n;
}
void ref_extractor_level_size(void* h, int level, int* lrows, int* lcols) {
    Probe* e = (Probe*)h;
    *lrows = e->mvImagePyramid[level].rows;
    *lcols = e->mvImagePyramid[level].cols;
}
void ref_extractor_level_image(void* h, int level, uint8_t* out) {
    const cv::Mat& m = ((Probe*)h)->mvImagePyramid[level];
    for (int r = 0; r < m.rows; r++) std::memcpy(out + (size_t)r * m.cols, m.ptr(r), (size_t)m.cols);
}
int ref_extractor_candidates(void* h, int level, int* xys, int cap) {
    const std::vector<int>& c = ((Probe*)h)->candidates[level];
    const int n = (int)c.size() / 3;
    if (xys && n <= cap) std::memcpy(xys, c.data(), c.size() * sizeof(int));
    return n;
}
// DistributeOctTree on a caller-supplied candidate list ((x, y, response) triples relative to (minX, minY))
int ref_distribute(void* h, const int* xys, int n, int minX, int maxX, int minY, int maxY, int N, int* out_xys, int cap) {
    Probe* e = (Probe*)h;
    std::vector<cv::KeyPoint> in;
    in.reserve(n);
    for (int i = 0; i < n; i++) in.push_back(cv::KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1.f, (float)xys[3 * i + 2]));
    std::vector<cv::KeyPoint> out = e->DistributeOctTree(in, minX, maxX, minY, maxY, N, 0);
    if ((int)out.size() <= cap)
        for (size_t i = 0; i < out.size(); i++) {
            out_xys[3 * i] = (int)out[i].pt.x; out_xys[3 * i + 1] = (int)out[i].pt.y; out_xys[3 * i + 2] = (int)out[i].response;
        }
    return (int)out.size();
}

// ---- Diasss::FEAmatcher ------------------------------------------------------------------------------------------
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return Diasss::FEAmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, (void*)a), cv::Mat(1, 32, CV_8U, (void*)b));
}

void ref_geo_nn_search(const orc_frame* f, const orc_frame* ref, int* corres_id, int* scc_count, double* scc_model,
                       int scc_cap, int* n_scc) {
    Diasss::Frame a = make_frame(f), b = make_frame(ref);
    std::vector<std::pair<int, double>> scc;
    std::vector<int> id = Diasss::FEAmatcher::GeoNearNeighSearch(a.img_id, b.img_id, a.norm_img, b.norm_img, a.kps, a.dst,
                                                                 a.geo_img, b.kps, b.dst, b.geo_img, scc);
    for (size_t i = 0; i < id.size(); i++) corres_id[i] = id[i];
    if (n_scc) *n_scc = (int)scc.size();
    for (size_t i = 0; i < scc.size() && (int)i < scc_cap; i++) { scc_count[i] = scc[i].first; scc_model[i] = scc[i].second; }
}

int ref_consistent_check(const orc_frame* s, const orc_frame* t, const int* corres1, const int* corres2, int scc1_count,
                         double scc1_model, int scc2_count, double scc2_model, orc_keypoint* src_keys,
                         orc_keypoint* tgt_keys, int cap) {
    Diasss::Frame a = make_frame(s), b = make_frame(t);
    std::vector<int> c1(corres1, corres1 + s->n), c2(corres2, corres2 + t->n);
    std::vector<std::pair<int, double>> s1, s2;
    if (scc1_count >= 0) s1.push_back(std::make_pair(scc1_count, scc1_model));
    if (scc2_count >= 0) s2.push_back(std::make_pair(scc2_count, scc2_model));
    std::vector<cv::KeyPoint> sk, tk;
    Diasss::FEAmatcher::ConsistentCheck(a, b, c1, c2, s1, s2, sk, tk);
    if ((int)sk.size() <= cap) { copy_out(sk, src_keys); copy_out(tk, tgt_keys); }
    return (int)sk.size();
}

// RobustMatching(Source, Target): rows6 = what was appended to Source.corres_kps; mirror = Target.corres_kps
int ref_robust_matching(const orc_frame* s, const orc_frame* t, double* rows6, double* mirror6, int cap) {
    Diasss::Frame a = make_frame(s), b = make_frame(t);
    Diasss::FEAmatcher::RobustMatching(a, b);
    const int k = a.corres_kps.rows;
    if (k <= cap)
        for (int i = 0; i < k; i++) {
            std::memcpy(rows6 + (size_t)i * 6, a.corres_kps.ptr(i), 48);
            if (mirror6) std::memcpy(mirror6 + (size_t)i * 6, b.corres_kps.ptr(i), 48);
        }
    return k;
}

float ref_compute_intersection(const double* sx, const double* sy, int srows, int scols, const double* tx, const double* ty,
                               int trows, int tcols) {
    std::vector<cv::Mat> s{wrap_f64(sx, srows, scols), wrap_f64(sy, srows, scols)};
    std::vector<cv::Mat> t{wrap_f64(tx, trows, tcols), wrap_f64(ty, trows, tcols)};
    return Diasss::Util::ComputeIntersection(s, t);
}

// ---- Diasss::Optimizer::GetKpsPairs(USE_ANNO = false, ...) (optimizer.cpp:575-639) on K x 6 corres_kps rows -------------
int ref_get_kps_pairs(const double* rows6, int k, int id_s, int id_t, const double* alt_s, int n_alt_s, const double* gra_s, int n_gra_s,
                      const double* alt_t, int n_alt_t, const double* gra_t, int n_gra_t, double* out7) {
    cv::Mat kps = k ? wrap_f64(rows6, k, 6) : cv::Mat();
    std::vector<gtsam::Vector7> r = Diasss::Optimizer::GetKpsPairs(false, kps, id_s, id_t, std::vector<double>(alt_s, alt_s + n_alt_s),
                                                                   std::vector<double>(gra_s, gra_s + n_gra_s),
                                                                   std::vector<double>(alt_t, alt_t + n_alt_t),
                                                                   std::vector<double>(gra_t, gra_t + n_gra_t));
    for (size_t i = 0; i < r.size(); i++)
        for (int c = 0; c < 7; c++) out7[i * 7 + c] = r[i](c);
    return (int)r.size();
}

// ---- Diasss::Frame (constructor = GetNormalizeSSS, GetFilteredMask, GetGeoImg, DetectFeature) ----------------------
void* ref_frame_create(int id, const double* raw, int rows, int cols, const double* pose6, const double* alt, int n_alt,
                       const double* g_range, int n_range) {
    cv::Mat img = wrap_f64(raw, rows, cols).clone(), pose = wrap_f64(pose6, rows, 6).clone(), anno;
    return new Diasss::Frame(id, img, pose, std::vector<double>(alt, alt + n_alt), std::vector<double>(g_range, g_range + n_range), anno);
}
void ref_frame_destroy(void* h) { delete (Diasss::Frame*)h; }
int ref_frame_nkps(void* h) { return (int)((Diasss::Frame*)h)->kps.size(); }
void ref_frame_get(void* h, uint8_t* norm_img, uint8_t* mask, double* geo_x, double* geo_y, orc_keypoint* kps, uint8_t* desc) {
    Diasss::Frame* f = (Diasss::Frame*)h;
    const int rows = f->norm_img.rows, cols = f->norm_img.cols;
    for (int r = 0; r < rows; r++) {
        if (norm_img) std::memcpy(norm_img + (size_t)r * cols, f->norm_img.ptr(r), (size_t)cols);
        if (mask) std::memcpy(mask + (size_t)r * cols, f->flt_mask.ptr(r), (size_t)cols);
        if (geo_x) std::memcpy(geo_x + (size_t)r * cols, f->geo_img[0].ptr(r), (size_t)cols * 8);
        if (geo_y) std::memcpy(geo_y + (size_t)r * cols, f->geo_img[1].ptr(r), (size_t)cols * 8);
    }
    if (kps) copy_out(f->kps, kps);
    if (desc) for (int i = 0; i < f->dst.rows; i++) std::memcpy(desc + (size_t)i * 32, f->dst.ptr(i), 32);
}
int ref_frame_corres_rows(void* h) { return ((Diasss::Frame*)h)->corres_kps.rows; }
void ref_frame_corres(void* h, double* rows6) {
    Diasss::Frame* f = (Diasss::Frame*)h;
    for (int i = 0; i < f->corres_kps.rows; i++) std::memcpy(rows6 + (size_t)i * 6, f->corres_kps.ptr(i), 48);
}
// RobustMatching on two constructed frames (appends to both frames' corres_kps, like the reference)
int ref_frame_robust_matching(void* s, void* t) {
    Diasss::Frame* a = (Diasss::Frame*)s;
    const int before = a->corres_kps.rows;
    Diasss::FEAmatcher::RobustMatching(*a, *(Diasss::Frame*)t);
    return a->corres_kps.rows - before;
}
float ref_frame_intersection(void* s, void* t) {
    return Diasss::Util::ComputeIntersection(((Diasss::Frame*)s)->geo_img, ((Diasss::Frame*)t)->geo_img);
}

// ---- test_demo's front end: src/diasss2.cpp:82-97 over host threads -----------------------------------------------
void* ref_survey_create(int threads) { Survey* s = new Survey; s->threads = std::max(1, threads); return s; }
void ref_survey_destroy(void* h) { delete (Survey*)h; }
// The buffers must stay valid until ref_survey_build returns.
void ref_survey_add(void* h, int id, const double* raw, int rows, int cols, const double* pose6, const double* alt,
                    int n_alt, const double* g_range, int n_range) {
    Survey* s = (Survey*)h;
    s->inputs.push_back(Survey::Input{id, rows, cols, raw, nullptr, nullptr, pose6, std::vector<double>(alt, alt + n_alt),
                                      std::vector<double>(g_range, g_range + n_range)});
}
// Same, for a frame given as its u8 norm_img and flt_mask (the benchmark's synthetic swaths)
void ref_survey_add_prepared(void* h, int id, const uint8_t* norm_img, const uint8_t* mask, int rows, int cols,
                             const double* pose6, const double* g_range, int n_range) {
    Survey* s = (Survey*)h;
    s->inputs.push_back(Survey::Input{id, rows, cols, nullptr, norm_img, mask, pose6, std::vector<double>(),
                                      std::vector<double>(g_range, g_range + n_range)});
}
// for i: Frame(i, img, pose, altitude, ground range, anno)        diasss2.cpp:82-84
void ref_survey_build(void* h) {
    Survey* s = (Survey*)h;
    s->frames.assign(s->inputs.size(), nullptr);
    parallel_for((int)s->inputs.size(), s->threads, [&](int i) {
        const Survey::Input& in = s->inputs[i];
        cv::Mat pose = wrap_f64(in.pose, in.rows, 6).clone(), anno;
        if (in.raw) {
            s->frames[i] = new Diasss::Frame(in.id, wrap_f64(in.raw, in.rows, in.cols).clone(), pose, in.alt, in.gr, anno);
        } else {
            s->frames[i] = new Diasss::Frame(blank_frame());
            prepare_from_images(*s->frames[i], in.id, cv::Mat(in.rows, in.cols, CV_8U, (void*)in.norm).clone(),
                                cv::Mat(in.rows, in.cols, CV_8U, (void*)in.mask).clone(), pose, in.gr);
        }
    });
    for (auto* f : s->frames) f->raw_img = cv::Mat();   // the front end never reads it again
    s->inputs.clear();
}
int ref_survey_nframes(void* h) { return (int)((Survey*)h)->frames.size(); }
void* ref_survey_frame(void* h, int i) { return ((Survey*)h)->frames[i]; }
// for i<j: if ComputeIntersection > min_overlap: RobustMatching    diasss2.cpp:88-97
// (all_pairs != 0 matches every i<j, BASELINE configs 3-4).  A pair works on copies of its two frames and the rows are
// appended to the frames' corres_kps afterwards in (i, j) order: the same rows in the same order as the serial loop.
int ref_survey_match(void* h, float min_overlap, int all_pairs) {
    Survey* s = (Survey*)h;
    const int F = (int)s->frames.size();
    s->pairs.clear();
    for (int i = 0; i < F; i++)
        for (int j = i + 1; j < F; j++) s->pairs.push_back(std::make_pair(i, j));
    const int P = (int)s->pairs.size();
    s->overlap.assign(P, 0.f);
    s->matched.assign(P, 0);
    s->pair_rows.assign(P, cv::Mat());
    std::vector<cv::Mat> mirror(P);
    parallel_for(P, s->threads, [&](int p) {
        const int i = s->pairs[p].first, j = s->pairs[p].second;
        const float ov = Diasss::Util::ComputeIntersection(s->frames[i]->geo_img, s->frames[j]->geo_img);
        s->overlap[p] = ov;
        if (!(all_pairs || ov > min_overlap)) return;
        Diasss::Frame a = *s->frames[i], b = *s->frames[j];
        a.corres_kps = cv::Mat();
        b.corres_kps = cv::Mat();
        Diasss::FEAmatcher::RobustMatching(a, b);
        s->matched[p] = 1;
        s->pair_rows[p] = a.corres_kps;
        mirror[p] = b.corres_kps;
    });
    int total = 0;
    for (int p = 0; p < P; p++) {
        if (!s->matched[p]) continue;
        s->frames[s->pairs[p].first]->corres_kps.push_back(s->pair_rows[p]);
        s->frames[s->pairs[p].second]->corres_kps.push_back(mirror[p]);
        total += s->pair_rows[p].rows;
    }
    return total;
}
// One bounded sample of the survey's work (bench.py's timed reference step): the listed frames are prepared again
// from their stored norm_img / flt_mask (GetGeoImg + DetectFeature, results discarded) and the listed pairs
// (indices into the i<j loop order) go through ComputeIntersection + RobustMatching on copies.  Returns the rows.
int ref_survey_sample_step(void* h, const int* frame_idx, int n_frames, const int* pair_idx, int n_pairs) {
    Survey* s = (Survey*)h;
    const int F = (int)s->frames.size();
    std::vector<std::pair<int, int>> all;
    for (int i = 0; i < F; i++)
        for (int j = i + 1; j < F; j++) all.push_back(std::make_pair(i, j));
    std::atomic<int> total{0};
    parallel_for(n_frames + n_pairs, s->threads, [&](int t) {
        if (t < n_frames) {
            const Diasss::Frame& src = *s->frames[frame_idx[t]];
            Diasss::Frame f = blank_frame();
            prepare_from_images(f, src.img_id, src.norm_img, src.flt_mask, src.dr_poses, src.ground_ranges);
        } else {
            const std::pair<int, int> p = all[pair_idx[t - n_frames]];
            const float ov = Diasss::Util::ComputeIntersection(s->frames[p.first]->geo_img, s->frames[p.second]->geo_img);
            (void)ov;
            Diasss::Frame a = *s->frames[p.first], b = *s->frames[p.second];
            a.corres_kps = cv::Mat();
            b.corres_kps = cv::Mat();
            Diasss::FEAmatcher::RobustMatching(a, b);
            total += a.corres_kps.rows;
        }
    });
    return total.load();
}
int ref_survey_threads(void* h) { return ((Survey*)h)->threads; }
int ref_survey_npairs(void* h) { return (int)((Survey*)h)->pairs.size(); }
void ref_survey_pair_info(void* h, float* overlap, int* matched, int* counts) {
    Survey* s = (Survey*)h;
    for (size_t p = 0; p < s->pairs.size(); p++) {
        overlap[p] = s->overlap[p];
        matched[p] = s->matched[p];
        counts[p] = s->pair_rows[p].rows;
    }
}
void ref_survey_rows(void* h, double* rows6) {   // all pairs' rows, concatenated in (i, j) order
    Survey* s = (Survey*)h;
    size_t o = 0;
    for (const cv::Mat& m : s->pair_rows)
        for (int i = 0; i < m.rows; i++, o++) std::memcpy(rows6 + o * 6, m.ptr(i), 48);
}

}  // extern "C"
