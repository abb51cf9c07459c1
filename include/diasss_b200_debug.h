/* diasss_b200 -- introspection entry points used by the stage-by-stage parity tests (tests/).
 * Not part of the drop-in boundary: they expose the intermediate products of the last extraction chunk so that
 * every reference stage can be compared with the oracle on its own. */
#ifndef DIASSS_B200_DEBUG_H
#define DIASSS_B200_DEBUG_H
#include "diasss_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Pyramid level `level` (>= 1) of image `image_in_chunk` of the last chunk: ComputePyramid, ORBextractor.cpp:1115-1140.
 * out: lrows x lcols bytes, tightly packed (host). */
int dsx_debug_level_image(dsx_ctx* ctx, int image_in_chunk, int level, uint8_t* out);

/* FAST candidates of one level in the reference's append order (ORBextractor.cpp:771-829): triples (x, y, response),
 * coordinates relative to (16,16).  *n = count (nothing is written when n > cap). */
int dsx_debug_candidates(dsx_ctx* ctx, int image_in_chunk, int level, int32_t* xys, int cap, int* n);

/* Keypoints selected by DistributeOctTree for one level, list order (ORBextractor.cpp:742-760 + :843-844):
 * triples (x, y, response) in level coordinates. */
int dsx_debug_level_keys(dsx_ctx* ctx, int image_in_chunk, int level, int32_t* xys, int cap, int* n);

/* RobustMatching with every intermediate: CorresID_1/2 after SCC, scc[0] = (inlier count, ModelX) per direction
 * (count 0 = empty scc), and the emitted rows / index pairs.  Any output may be NULL. */
int dsx_debug_match(dsx_ctx* ctx, const dsx_frame* source, const dsx_frame* target, int32_t* corres1, int32_t* corres2,
                    int32_t* scc_count2, double* scc_model2, double* rows6, int32_t* src_idx, int32_t* tgt_idx, int cap, int* k);

/* Phase clocks of fast_cells_kernel, only in a library built with -DDSX_FAST_PROFILE (tools/build_variant.sh):
 * out16[2p] = clocks summed over warps up to phase boundary p, out16[2p+1] = warps counted.  Phases: 0 staging issue /
 * own wait, 1 barrier, 2 shifted copies, 3 barrier, 4 scoring, 5 barrier, 6 per-cell listing.  DSX_ERR_INVALID otherwise. */
int dsx_debug_fast_profile(uint64_t* out16, int reset);

/* The device's cosf / sinf of the descriptor rotation (ORBextractor.cpp:113; glibc 2.39's algorithm) on n host
 * floats: s[i] = sinf(x[i]), c[i] = cosf(x[i]).  Compared with the host libm by tests/test_gpu_extract.py. */
int dsx_debug_sincosf(dsx_ctx* ctx, const float* x, float* s, float* c, int n);

#ifdef __cplusplus
}
#endif
#endif
