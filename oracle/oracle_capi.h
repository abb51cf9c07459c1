/* TEST INFRASTRUCTURE ONLY -- C interface of the CPU oracle (see orb_oracle.cpp header).
 * Loaded through ctypes by oracle/oracle.py; never by the product. */
#ifndef DSX_ORACLE_CAPI_H
#define DSX_ORACLE_CAPI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, size, angle, response; int octave, class_id; } orc_keypoint; /* cv::KeyPoint, 28 B */

/* Diasss::Frame fields read by the matcher (frame.h:30-46). geo_x/geo_y are the two
 * rows x cols CV_64F planes of Frame::geo_img. */
typedef struct {
    int img_id, rows, cols, n;
    const orc_keypoint* kps;
    const uint8_t* desc; /* n x 32 */
    const double* geo_x;
    const double* geo_y;
} orc_frame;

/* OpenCV primitives */
void orc_resize_linear_u8(const uint8_t* src, int srows, int scols, int sstep, uint8_t* dst, int drows, int dcols, int dstep);
int orc_fast9_16(const uint8_t* img, int rows, int cols, int step, int threshold, int* xys, int cap);
void orc_gaussian13_s2(const uint8_t* src, int rows, int cols, int sstep, uint8_t* dst, int dstep);
float orc_fast_atan2(float y, float x);
void orc_sincosf(float x, float* s, float* c);   /* the platform's cosf / sinf (glibc 2.39 algorithm) */
long orc_sincosf_scan(uint32_t first_bits, uint32_t count);
void orc_pattern(signed char* out1024);
void orc_rng_draws(uint32_t* out, int n);

/* ORBextractor */
void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orc_extractor_destroy(void* h);
void orc_extractor_tables(void* h, float* scale, float* inv_scale, int* features_per_level, int* umax16);
int orc_extractor_run(void* h, const uint8_t* img, int rows, int cols, int step, orc_keypoint* kps, uint8_t* desc, int cap);
void orc_extractor_level_size(void* h, int rows, int cols, int level, int* lrows, int* lcols);
void orc_extractor_level_image(void* h, int level, uint8_t* out);
int orc_extractor_candidates(void* h, int level, int* xys, int cap);
int orc_extractor_level_keys(void* h, int level, orc_keypoint* out, int cap);
int orc_distribute(const int* xys, int n, int minX, int maxX, int minY, int maxY, int N, int* out_xys);

/* Frame (frame.cpp) */
int orc_mask_filter(const orc_keypoint* kps, const uint8_t* desc, int n, const uint8_t* mask, int mstep,
                    orc_keypoint* out_kps, uint8_t* out_desc, int* out_index);
void orc_geo_img(int rows, int cols, const double* pose6, const double* g_range, int n_range, double* geo_x, double* geo_y);
double orc_mean(const double* raw, int rows, int cols, int order /*0 sequential, 1 the library's 32-lane order*/);
void orc_normalize_sss_m(const double* raw, int rows, int cols, double mean, uint8_t* out);
void orc_filtered_mask_m(const double* raw, int rows, int cols, double mean, uint8_t* out);
void orc_normalize_sss(const double* raw, int rows, int cols, uint8_t* out);
void orc_filtered_mask(const double* raw, int rows, int cols, uint8_t* out);
float orc_compute_intersection(const double* sx, const double* sy, int sn, const double* tx, const double* ty, int tn);

/* Optimizer::GetKpsPairs, USE_ANNO = 0 (optimizer.cpp:575-639); out7 holds up to k x 7 doubles; returns the count */
int orc_get_kps_pairs(const double* rows6, int k, int id_t, const double* alt_s, const double* gra_s, int n_gra_s, const double* alt_t,
                      const double* gra_t, int n_gra_t, double* out7);

/* FEAmatcher (ORB mode) */
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
/* One direction.  corres_id[n] final (after SCC); pre_corres[n] before SCC; best/sec/ncand per keypoint
 * (best,sec = 1000 when no candidate).  scc_* = the (inlier count, ModelX) pushes in order.  Any out pointer may be NULL. */
void orc_geo_nn_search(const orc_frame* f, const orc_frame* ref, int* corres_id, int* pre_corres, int* best_dist,
                       int* sec_dist, int* n_cand, int* scc_count, double* scc_model, int scc_cap, int* n_scc);
/* RobustMatching: returns K; rows6 = K x 6 doubles as appended to Source.corres_kps (FEAmatcher.cpp:37-40);
 * src_idx/tgt_idx = keypoint indices of each row. */
int orc_robust_matching(const orc_frame* s, const orc_frame* t, double* rows6, int* src_idx, int* tgt_idx, int cap,
                        int* corres1, int* corres2);

#ifdef __cplusplus
}
#endif
#endif
