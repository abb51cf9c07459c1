// K2 -- FAST-9/16 detection per ~30 px cell with in-cell 3x3 non-max suppression and the 12 -> 7
// per-cell threshold fallback.  Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree
// (ORBextractor.cpp:771-829) and the cv::FAST(roi, kps, th, true) calls inside it (:809, :814;
// OpenCV features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>).
//
// Semantics reproduced:
//   * m(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min d_k, min -d_k), d_k = I(p)-I(ring k);
//     corner at threshold t <=> m > t; score = m-1 (independent of t).
//   * scores exist only for ROI pixels x in [3,w-3), y in [3,h-3); everything else counts as 0, so NMS never
//     crosses a cell.
//   * keypoint <=> score >= t and score > all 8 neighbours; a cell falls back to minThFAST iff it has no
//     keypoint at iniThFAST.  NMS against the score map of the lower threshold is identical (sub-threshold
//     neighbours are smaller than any score >= t).
//   * output order inside a cell is row-major; coordinates are relative to the ROI origin.
//
// On textured sonar imagery 15-25 % of all pixels are corners at t = 7 and no cheap rejection test removes more
// than half of the rest, so the score is computed DENSELY, branch-free, two pixels per 32-bit operation:
//   m = max( max_k min_{arc k}(ring) - I(p),  I(p) - min_k max_{arc k}(ring) )
// needs only min/max over raw ring values.  Each 16-bit lane carries one pixel's ring value as the fp16 number
// 1024 + pixel (bits 0x6400 | pixel): positive, so the lanes order the same as fp16 numbers and as unsigned integers.
// The 9-arc extrema come from a 36-operation min/max network per polarity (see arc_pair_h); the 16 (min, max) pairs
// at its inputs run on the FMA pipe (HFMA2.RELU + 2 HADD2 per pair, exact on these integers), the other 40 operations
// are VIMNMX / VIMNMX3 .U16x2 on the ALU pipe, so that neither pipe alone carries the network.
//
// Mapping: one CTA = 8 warps = 8 horizontally adjacent cells of one cell row.
//   staging  the (hCell+6) x SP byte strip arrives as ONE tensor-map box (TMA); fallbacks: a bulk copy per row, or
//            4-byte cp.async copies when the planes are not 16-byte aligned.
//   phase A0 all 256 threads expand the strip into two planes of 16-bit lanes (pixel pairs starting at even / odd x),
//            so that every ring window of a pixel pair is one aligned 4-byte shared-memory load.
//   phase A  all 256 threads: one pixel pair per thread and step, neighbouring lanes on neighbouring pairs, each thread
//            walking down consecutive rows (the ring's +-3 columns carry over); scores go to a shared score map with
//            zero border rows, which takes the raw strip's place.
//   phase B  one warp per cell, warp-synchronous: scan the cell's scores (row-major, ballot compaction) into a
//            corner list; NMS per corner (neighbours outside the cell's detection rectangle count as 0);
//            threshold decision by ballot; ordered emission into the cell's staging slot + count.
// (DSX_FAST_FP16=0 builds the round-1 form: byte lanes in the high byte of 16-bit lanes over byte-shifted strip copies,
// the whole network on the ALU pipe -- 11.7 ms instead of 9.95 ms on BASELINE config 4.)
// Bound: instruction issue (82 % of the issue slots; ALU pipe 64 %, FMA pipe 59 % busy -- DESIGN.md section 4); HBM traffic is one
// read of every level.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_pipeline.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "dsx_internal.cuh"

namespace dsx {

namespace {

#ifndef DSX_FAST_WARPS
#define DSX_FAST_WARPS 8
#endif
constexpr int kWarps = DSX_FAST_WARPS;          // cells per strip = warps per CTA
constexpr int kLX = 8 * kWarps;                  // phase A: threads across a strip row (one aligned word each)
constexpr int kLXs = (kWarps == 8) ? 6 : (kWarps == 4) ? 5 : 4;   // log2(kLX)
static_assert((1 << kLXs) == kLX, "kWarps must be 2, 4 or 8");
constexpr int kPad = 4;   // bytes of addressable padding in front of every strip / score row

// DSX_FAST_FP16 = 1 (default): phase A works on 16-bit lanes holding 1024 + pixel as an fp16 number (bit pattern
// 0x6400 | pixel).  All values are positive, so their order as fp16 numbers is their order as unsigned 16-bit integers:
// part of the min/max network then runs as HADD2 / HFMA2.RELU on the FMA pipe (max(a,b) = b + relu(a - b),
// min(a,b) = a - relu(a - b), exact on integers below 2048) while the rest stays VIMNMX on the ALU pipe -- the two pipes
// take one instruction every other clock each, and the all-integer network left the FMA pipe idle.
// DSX_FAST_FP16 = 0: the round-1 all-integer network on byte-shifted strip copies.
#ifndef DSX_FAST_FP16
#define DSX_FAST_FP16 1
#endif
// how many of the 8 end-point extrema pairs go to the FMA pipe (the 8 pair extrema always do).  Measured on config 4:
// 8 -> 9.95 ms, 4 -> 10.18 ms, 2 -> 10.06 ms (profiles/r02/README.md)
#ifndef DSX_FAST_END_FMA
#define DSX_FAST_END_FMA 8
#endif

__device__ __forceinline__ uint32_t ld4(const uint8_t* s, int off) {
    // unaligned 4-byte window from shared memory: two aligned words + funnel shift
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (off & ~3));
    return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}

// 4-byte window starting at byte S (0..8) of the 12 bytes (P,Q,N)
template <int S>
__device__ __forceinline__ uint32_t win(uint32_t P, uint32_t Q, uint32_t N) {
    if (S == 0) return P;
    if (S < 4) return __funnelshift_r(P, Q, 8 * S);
    if (S == 4) return Q;
    if (S < 8) return __funnelshift_r(Q, N, 8 * (S - 4));
    return N;
}

#if !DSX_FAST_FP16
// arc measure of two pixels whose ring values sit in the high bytes of the 16-bit lanes of r[0..15]; c = centres.
// returns max(m, t_floor) - 1 per pixel in the low bytes of the two 16-bit lanes, m = max(a - c, c - b) (0x8000-biased
// lanes keep the tail packed).  With t_floor = max(min threshold, 1) every corner at either threshold keeps its score
// m - 1 >= t_floor and every other pixel reads t_floor - 1: the score map has a FLOOR instead of zeros.  That is
// equivalent for everything downstream -- non-maximum suppression only asks whether a corner's score exceeds its
// neighbours', floor pixels never exceed a neighbour inside the detection area (all >= floor), and emission requires
// score >= threshold > floor -- and it saves the select that would zero them.
__device__ __forceinline__ uint32_t arc_pair(const uint32_t (&r)[16], uint32_t c, uint32_t t_floor) {
    // The arcs starting at 2j and 2j+1 share the eight ring values 2j+1 .. 2j+8 (the pairing of OpenCV's cornerScore<16>):
    //   max(min(arc 2j), min(arc 2j+1)) = min( min r[2j+1 .. 2j+8], max(r[2j], r[2j+9]) )
    // so per polarity: 8 pair extrema at the odd starts, 8 four-blocks, 8 end-point extrema, 8 three-input joins and a
    // 4-operation reduction = 36 operations (the two-level three-input network over all 16 arcs costs 40).
    uint32_t pn[8], px[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        pn[j] = __vminu2(r[2 * j + 1], r[(2 * j + 2) & 15]);
        px[j] = __vmaxu2(r[2 * j + 1], r[(2 * j + 2) & 15]);
    }
    uint32_t qn[8], qx[8];      // extrema of r[2j+1 .. 2j+4]
#pragma unroll
    for (int j = 0; j < 8; j++) {
        qn[j] = __vminu2(pn[j], pn[(j + 1) & 7]);
        qx[j] = __vmaxu2(px[j], px[(j + 1) & 7]);
    }
    uint32_t wn[8], wx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t en = __vmaxu2(r[2 * j], r[(2 * j + 9) & 15]), ex = __vminu2(r[2 * j], r[(2 * j + 9) & 15]);
        wn[j] = __vimin3_u16x2(qn[j], qn[(j + 2) & 7], en);
        wx[j] = __vimax3_u16x2(qx[j], qx[(j + 2) & 7], ex);
    }
    uint32_t a = __vimax3_u16x2(__vimax3_u16x2(wn[0], wn[1], wn[2]), __vimax3_u16x2(wn[3], wn[4], wn[5]), __vmaxu2(wn[6], wn[7]));
    uint32_t b = __vimin3_u16x2(__vimin3_u16x2(wx[0], wx[1], wx[2]), __vimin3_u16x2(wx[3], wx[4], wx[5]), __vminu2(wx[6], wx[7]));
    const uint32_t A = __byte_perm(a, 0, 0x4341), B = __byte_perm(b, 0, 0x4341), C = __byte_perm(c, 0, 0x4341);   // 0x00vv per lane
    const uint32_t X = A + 0x80008000u - C, Y = C + 0x80008000u - B;          // a-c and c-b, biased; no lane borrows
    return __vimax3_u16x2(X, Y, 0x80008000u + t_floor * 0x00010001u) - 0x80018001u;   // max(m, t_floor) - 1 per lane
}

#endif

// (a, b) -> (min, max) of two fp16x2 words holding integers in [1024, 1279], on the FMA pipe: 3 instructions for both
__device__ __forceinline__ void hminmax(uint32_t a, uint32_t b, uint32_t& mn, uint32_t& mx) {
    const __half2 ha = *reinterpret_cast<const __half2*>(&a), hb = *reinterpret_cast<const __half2*>(&b);
    const __half2 d = __hfma2_relu(hb, __float2half2_rn(-1.0f), ha);      // relu(a - b)
    const __half2 hx = __hadd2(hb, d), hn = __hsub2(ha, d);
    mx = *reinterpret_cast<const uint32_t*>(&hx);
    mn = *reinterpret_cast<const uint32_t*>(&hn);
}

// arc measure of the two pixels of a pair; r[0..15] = ring values, c = centres, each lane an fp16 number 1024 + pixel.
// Same network as arc_pair.  floor_bits = fp16x2 bits of t_floor + 256.  Returns per lane the fp16 number
// 1024 + max(m, t_floor) - 1, i.e. the score byte in the low byte of the lane.
__device__ __forceinline__ uint32_t arc_pair_h(const uint32_t (&r)[16], uint32_t c, uint32_t floor_bits) {
    uint32_t pn[8], px[8];
#pragma unroll
    for (int j = 0; j < 8; j++) hminmax(r[2 * j + 1], r[(2 * j + 2) & 15], pn[j], px[j]);
    uint32_t qn[8], qx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        qn[j] = __vminu2(pn[j], pn[(j + 1) & 7]);
        qx[j] = __vmaxu2(px[j], px[(j + 1) & 7]);
    }
    uint32_t wn[8], wx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t en, ex;
        if (j < DSX_FAST_END_FMA) hminmax(r[2 * j], r[(2 * j + 9) & 15], ex, en);
        else { en = __vmaxu2(r[2 * j], r[(2 * j + 9) & 15]); ex = __vminu2(r[2 * j], r[(2 * j + 9) & 15]); }
        wn[j] = __vimin3_u16x2(qn[j], qn[(j + 2) & 7], en);
        wx[j] = __vimax3_u16x2(qx[j], qx[(j + 2) & 7], ex);
    }
    const uint32_t a = __vimax3_u16x2(__vimax3_u16x2(wn[0], wn[1], wn[2]), __vimax3_u16x2(wn[3], wn[4], wn[5]), __vmaxu2(wn[6], wn[7]));
    const uint32_t b = __vimin3_u16x2(__vimin3_u16x2(wx[0], wx[1], wx[2]), __vimin3_u16x2(wx[3], wx[4], wx[5]), __vminu2(wx[6], wx[7]));
    // X = a - c + 256 and Y = c - b + 256 are integers in [1, 511]: exact, positive, ordered like their bits
    const __half2 ha = *reinterpret_cast<const __half2*>(&a), hb = *reinterpret_cast<const __half2*>(&b);
    const __half2 hc = *reinterpret_cast<const __half2*>(&c);
    const __half2 k256 = __float2half2_rn(256.0f);
    const __half2 X = __hadd2(ha, __hsub2(k256, hc)), Y = __hsub2(__hadd2(hc, k256), hb);
    const uint32_t m = __vimax3_u16x2(*reinterpret_cast<const uint32_t*>(&X), *reinterpret_cast<const uint32_t*>(&Y), floor_bits);
    const __half2 sc = __hadd2(*reinterpret_cast<const __half2*>(&m), __float2half2_rn(767.0f));     // + 1024 - 257
    return *reinterpret_cast<const uint32_t*>(&sc);
}

// The score map's rows 0 and hd+1 (the neighbours above the first and below the last detection row) must read 0; every
// other byte phase B looks at is written by phase A (detection columns) or masked off (columns of another cell).
__device__ __forceinline__ void zero_score_border(uint8_t* score, int SP, int hd, int tid) {
    const int W = SP >> 2;
    uint32_t* sc = reinterpret_cast<uint32_t*>(score);
    for (int i = tid; i < 2 * W; i += kWarps * 32) sc[i < W ? i : (hd + 1) * W + (i - W)] = 0;
}

// DSX_FAST_PROFILE (tools/build_variant.sh prof -DDSX_FAST_PROFILE): every warp adds the clocks it spent up to each
// phase boundary into A.prof[phase * 2] and counts itself in A.prof[phase * 2 + 1]; read with dsx_debug_fast_profile.
#ifdef DSX_FAST_PROFILE
__device__ unsigned long long g_fast_prof[512][16];     // spread over 512 slots: the marks must not contend
#define FAST_MARK(ph) do { if ((threadIdx.x & 31) == 0) { const long long t_ = clock64(); unsigned long long* s_ = g_fast_prof[(blockIdx.x * 8 + (threadIdx.x >> 5) + blockIdx.y * 37) & 511]; atomicAdd(s_ + 2 * (ph), (unsigned long long)(t_ - t_prev)); atomicAdd(s_ + 2 * (ph) + 1, 1ull); t_prev = t_; } } while (0)
#else
#define FAST_MARK(ph) do { } while (0)
#endif

struct FastArgs {
    LevelGeom g;
    const uint8_t* img;      // level plane of image 0
    long long img_stride;    // bytes between images
    int pitch;               // row pitch of the level plane
    int32_t* cell_count;     // [n][cells_total]
    uint32_t* stage;         // [n][stage_total]
    long long cells_total, stage_total;
    int ini_th, min_th;
    int SP;                  // strip / score-map pitch (bytes, multiple of 4)
    int strip_bytes, score_bytes, list_bytes;
    // K3's count grid: every emitted key is histogrammed into its depth-D cell of DivideNode's fixed grid
    int32_t* hist;           // [n][hist_total]
    unsigned long long* gbest; // [n][hist_total]  best key per grid cell: response << 56 | (kBestOrderMask - order)
    long long hist_total;
    const uint16_t* xlut;    // [w]  root << 8 | depth-D column     (this level)
    const uint8_t* ylut;     // [h]  depth-D row
    int use_tma;             // 1: the strip's rows arrive as bulk asynchronous copies (TMA unit, cp.async.bulk), one per row
                             // 2: the whole strip arrives as ONE tensor-map box (cp.async.bulk.tensor.3d, UTMALDG)
    CUtensorMap tmap;        // use_tma == 2: {x in 32-bit words, y, image} over this level's planes
};

// One launch covers every level of the chunk: blocks [first[l], first[l + 1]) work on level l, largest level first, so
// that the small levels fill the tail of the large one instead of each launch draining the GPU on its own (a chunk of
// the host pipeline is only 4 images; six launches per chunk cost it ~1 ms per survey in drained tails).
constexpr int kFastMaxLevels = 8;
struct FastArgsAll {
    FastArgs lv[kFastMaxLevels];
    int first[kFastMaxLevels + 1];
    int nlevels;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SPC: the strip pitch as a compile-time constant (0 = read it from the arguments); with it every ring window of phase A is
// a load at an immediate offset from one pointer per plane instead of an address computed per window.
template <int SPC>
__global__ void __launch_bounds__(kWarps * 32, 32 / kWarps) fast_cells_kernel(const __grid_constant__ FastArgsAll AA) {
    int level = 0;
    while (level + 1 < AA.nlevels && (int)blockIdx.x >= AA.first[level + 1]) level++;
    const FastArgs& A = AA.lv[level];
    const int bx = (int)blockIdx.x - AA.first[level];        // strip of this level
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long tma_bar;
#ifdef DSX_FAST_PROFILE
    long long t_prev = clock64();
#endif
    const LevelGeom& g = A.g;
    const int groupsX = (g.nCols + kWarps - 1) / kWarps;
    const int ci = bx / groupsX;                   // cell row
    const int j0 = (bx % groupsX) * kWarps;
    const int iniY = kMinBorder + ci * g.hCell;
    if (iniY >= g.maxBY - 3) return;               // ORBextractor.cpp:794
    const int maxY = min(iniY + g.hCell + 6, g.maxBY);
    const int hROI = maxY - iniY, hd = hROI - 6;
    // cells of this strip that the reference does not skip (:803); the skip test is monotone in j
    int ncell = min(kWarps, g.nCols - j0);
    while (ncell > 0 && kMinBorder + (j0 + ncell - 1) * g.wCell >= g.maxBX - 6) ncell--;
    if (ncell == 0 || hd <= 0) return;             // counts stay 0 (memset by the launcher)
    const int xBegin = kMinBorder + j0 * g.wCell;
    const int xEnd = min(kMinBorder + (j0 + ncell) * g.wCell + 6, g.maxBX);
    const int xa = xBegin & ~3;
    // strip column of global x = x - xorg.  The bulk-copy path needs 16-byte aligned row segments on both sides, so
    // its origin is rounded down to 16 (the fallback keeps kPad bytes in front of xa)
    const int xorg = A.use_tma ? ((xa - kPad) & ~15) : (xa - kPad);
    const int SP = SPC ? SPC : A.SP;
    uint8_t* strip = smem;                         // [hROI][SP], column of global x = x - xorg
    // three more copies of the strip follow it, copy k shifted by k bytes (word w of copy k = bytes 4w+k .. 4w+k+3 of the
    // row), so that every unaligned 4-byte ring window of phase A is ONE aligned shared-memory load instead of two
    // loads and a funnel shift on the saturated integer pipe
#if DSX_FAST_FP16
    // fp16 lanes: the raw strip is followed by two planes of 16-bit lanes (0x6400 | pixel), plane E holding the pixel
    // pairs (2j, 2j+1) and plane O the pairs (2j+1, 2j+2) as one word each, so that every ring window of a pixel pair
    // is ONE aligned load.  The score map [hd + 2][SP] takes the raw strip's place once the planes are written.
    uint8_t* score = smem;
    uint32_t* planeE = reinterpret_cast<uint32_t*>(smem + A.strip_bytes);
    uint32_t* planeO = planeE + (A.strip_bytes >> 1);
#else
    uint8_t* score = smem + 4 * A.strip_bytes;     // [hd + 2][SP], same column mapping, rows shifted by one
#endif
    const uint8_t* img = A.img + (long long)blockIdx.y * A.img_stride;
    const int tid = threadIdx.x;

    // ---- stage the strip, rows iniY..maxY; zero the score map's border rows meanwhile.
    //      TMA path: one bulk asynchronous copy (cp.async.bulk, the TMA unit) per strip row, issued by the lanes of
    //      warp 0, all completing on one mbarrier -- ~36 instructions per CTA instead of ~2400 4-byte copies.
    //      Fallback (pitch / base not 16-byte aligned): 4-byte cp.async copies issued by all threads.
    if (A.use_tma == 2) {
        // ONE tensor-map box {SP / 4 words, hROI rows, this image}: the TMA unit generates the row requests itself.  The
        // box starts at byte xorg of the row -- a multiple of 16, which the unit requires of the innermost coordinate
        // (a start at byte 12 raises a fault that is reported as "illegal instruction": tools/ubench/tma_test.cu).
        // Rows / words beyond the plane are zero-filled by the unit and count as transferred.
        const uint32_t bar = smem_u32(&tma_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SP * (g.hCell + 6))) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(strip)), "l"(&A.tmap), "r"(xorg >> 2), "r"(iniY), "r"((int)blockIdx.y), "r"(bar) : "memory");
        }
#if !DSX_FAST_FP16
        zero_score_border(score, SP, hd, tid);
#endif
        __syncthreads();           // (the barrier word is initialised before anybody polls it)
        if (tid < 32) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
    } else if (A.use_tma) {
        const uint32_t bar = smem_u32(&tma_bar);
        const uint32_t row_bytes = (uint32_t)min(SP, A.pitch - xorg);       // multiple of 16, stays inside the row pitch
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t)hROI) : "memory");
        }
        __syncthreads();
        if (tid < 32)
            for (int r = tid; r < hROI; r += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(strip + r * SP)), "l"(img + (long long)(iniY + r) * A.pitch + xorg), "r"(row_bytes), "r"(bar) : "memory");
#if !DSX_FAST_FP16
        zero_score_border(score, SP, hd, tid);
#endif
        if (tid < 32) {            // one warp waits for the bytes; the others sleep at the barrier below instead of polling
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
    } else {
        const int nwords = (xEnd - xa + 3) >> 2;
        const int total = nwords * hROI;
        const int step_r = (kWarps * 32) / nwords, step_w = (kWarps * 32) - step_r * nwords;
        int r = tid / nwords, w = tid - r * nwords;
        for (int e = tid; e < total; e += kWarps * 32) {
            __pipeline_memcpy_async(strip + r * SP + (xa - xorg) + 4 * w, img + (long long)(iniY + r) * A.pitch + xa + 4 * w, 4);
            r += step_r; w += step_w;
            if (w >= nwords) { w -= nwords; r++; }
        }
        __pipeline_commit();
#if !DSX_FAST_FP16
        zero_score_border(score, SP, hd, tid);
#endif
        __pipeline_wait_prior(0);
    }
    FAST_MARK(0);          // staging issued / own wait done
    __syncthreads();
    FAST_MARK(1);          // barrier after staging

#if DSX_FAST_FP16
    // ---- phase A0: the two planes of fp16 lanes (linear over the strip's words; a row's last O word borrows the next
    //      row's first byte, which no window reads)
    {
        const uint32_t* s0 = reinterpret_cast<const uint32_t*>(strip);
        const int total = (SP >> 2) * hROI;
        const uint32_t K = 0x64646464u;
        for (int e = tid; e < total; e += kWarps * 32) {
            const uint32_t a = s0[e], b = s0[e + 1];
            uint2 E, O;
            E.x = __byte_perm(a, K, 0x4140); E.y = __byte_perm(a, K, 0x4342);
            O.x = __byte_perm(a, K, 0x4241); O.y = __byte_perm(__byte_perm(a, b, 0x0043), K, 0x4140);
            *reinterpret_cast<uint2*>(planeE + 2 * e) = E;
            *reinterpret_cast<uint2*>(planeO + 2 * e) = O;
        }
    }
    FAST_MARK(2);          // planes
    __syncthreads();
    FAST_MARK(3);          // barrier after the planes
    zero_score_border(score, SP, hd, tid);       // (the raw strip is dead; phase B reads these rows after the next barrier)

    // ---- phase A: dense scores, one pixel pair per thread and iteration; neighbouring lanes take neighbouring pairs
    //      (conflict-free 4-byte loads), a thread keeps its pair column and walks down the rows
    const uint32_t t_lo = (uint32_t)max(min(A.ini_th, A.min_th), 1);   // the score map's floor is t_lo - 1 (see arc_pair)
    {
        constexpr int kPX = kWarps * 16;                          // pair columns per sweep; the CTA covers 2 rows per sweep
        const int d0 = xBegin + 3 - xorg, d1 = xEnd - 3 - xorg;   // detection columns (strip coordinates)
        const int p0 = d0 >> 1, npx = ((d1 + 1) >> 1) - p0;
        const int W2 = SP >> 1;                                   // plane words per row
        const __half2 fl = __float2half2_rn((float)(t_lo + 256));
        const uint32_t floor_bits = *reinterpret_cast<const uint32_t*>(&fl);
        for (int px = tid & (kPX - 1); px < npx; px += kPX) {
            const int e = p0 + px;                                // pair index: pixels 2e, 2e + 1
            uint32_t keep = 0xffffu;                              // bytes outside the detection columns stay 0
            if (2 * e < d0) keep &= 0xff00u;
            if (2 * e + 1 >= d1) keep &= 0x00ffu;
            // a thread walks down CONSECUTIVE rows (the upper or the lower half of the strip): the ring's columns dx = +-3
            // span the rows -1, 0, +1, so two of their three words per side carry over from the row before
            const int half = (hd + 1) >> 1;
            const int rbeg = (tid / kPX) ? half : 0, rend = (tid / kPX) ? hd : half;
            const uint32_t* qe = planeE + (rbeg + 3) * W2 + e;
            const uint32_t* qo = planeO + (rbeg + 3) * W2 + e;
            uint16_t* o = reinterpret_cast<uint16_t*>(score + (rbeg + 1) * SP) + e;
            uint32_t Rm = qo[-W2 + 1], Rc = qo[1], Lm = qo[-W2 - 2], Lc = qo[-2];
#pragma unroll 3
            for (int row = rbeg; row < rend; row++, qe += W2, qo += W2, o += SP >> 1) {
                uint32_t r[16];
                const uint32_t Rp = qo[W2 + 1], Lp = qo[W2 - 2];
                // ring window (dx, dy) of the pair: pixels 2e + dx, 2e + dx + 1 = word e + dx / 2 of plane E (dx even) or
                // word e + (dx - 1) / 2 of plane O (dx odd)
#define RING(k, dx, dy) r[k] = ((dx) & 1) ? qo[(dy) * W2 + (((dx) - 1) >> 1)] : qe[(dy) * W2 + ((dx) >> 1)];
                RING(0, 0, 3)   RING(1, 1, 3)   RING(2, 2, 2)    r[3] = Rp;
                r[4] = Rc;      r[5] = Rm;      RING(6, 2, -2)   RING(7, 1, -3)
                RING(8, 0, -3)  RING(9, -1, -3) RING(10, -2, -2) r[11] = Lm;
                r[12] = Lc;     r[13] = Lp;     RING(14, -2, 2)  RING(15, -1, 3)
#undef RING
                const uint32_t sc = arc_pair_h(r, qe[0], floor_bits);
                *o = (uint16_t)(__byte_perm(sc, 0, 0x4420) & keep);
                Rm = Rc; Rc = Rp; Lm = Lc; Lc = Lp;
            }
        }
    }
#else
    // ---- phase A0: the byte-shifted copies (linear over the strip's words; a row's last word borrows the next row's
    //      first bytes, which no window reads)
    {
        const uint32_t* s0 = reinterpret_cast<const uint32_t*>(strip);
        const int cw = A.strip_bytes >> 2;                              // words per copy
        uint32_t* s1 = reinterpret_cast<uint32_t*>(strip) + cw;
        const int total = (SP >> 2) * hROI;
        for (int e = tid; e < total; e += kWarps * 32) {
            const uint32_t a = s0[e], b = s0[e + 1];
            s1[e] = __funnelshift_r(a, b, 8);
            s1[e + cw] = __funnelshift_r(a, b, 16);
            s1[e + 2 * cw] = __funnelshift_r(a, b, 24);
        }
    }
    FAST_MARK(2);          // shifted copies
    __syncthreads();
    FAST_MARK(3);          // barrier after the copies

    // ---- phase A: dense scores, one aligned 4-pixel word per thread and iteration; a thread keeps its column and
    //      walks down the rows (4 rows per sweep of the CTA)
    const uint32_t t_lo = (uint32_t)max(min(A.ini_th, A.min_th), 1);   // the score map's floor is t_lo - 1 (see arc_pair)
    {
        const int d0 = xBegin + 3 - xorg, d1 = xEnd - 3 - xorg;   // detection columns (strip coordinates)
        const int w0 = d0 >> 2, nwx = ((d1 + 3) >> 2) - w0;
        const int W = SP >> 2, CW = A.strip_bytes >> 2;
        for (int wx = tid & (kLX - 1); wx < nwx; wx += kLX) {
            const int col = (w0 + wx) << 2;
            // bytes outside the detection columns stay 0
            uint32_t keep = 0xffffffffu;
            if (d0 - col > 0) keep &= 0xffffffffu << (8 * (d0 - col));
            if (d1 - col < 4) keep &= 0xffffffffu >> (8 * (4 - (d1 - col)));
            const uint32_t* q = reinterpret_cast<const uint32_t*>(strip + ((tid >> kLXs) + 3) * SP + col);
            uint32_t* o = reinterpret_cast<uint32_t*>(score + ((tid >> kLXs) + 1) * SP + col);
            for (int row = tid >> kLXs; row < hd; row += (kWarps * 32) >> kLXs, q += 4 * W, o += 4 * W) {
                uint32_t ro[16], re[16];   // ring windows for pixels (1,3) and (0,2)
                // window starting S bytes after the word in front of q = word (S >> 2) - 1 of byte-shifted copy S & 3
#define WIN(S, dy) q[((S) & 3) * CW + (dy) * W + ((S) >> 2) - 1]
#define RING(k, dx, dy) { ro[k] = WIN(4 + (dx), dy); re[k] = WIN(3 + (dx), dy); }
                RING(0, 0, 3)   RING(1, 1, 3)   RING(2, 2, 2)    RING(3, 3, 1)
                RING(4, 3, 0)   RING(5, 3, -1)  RING(6, 2, -2)   RING(7, 1, -3)
                RING(8, 0, -3)  RING(9, -1, -3) RING(10, -2, -2) RING(11, -3, -1)
                RING(12, -3, 0) RING(13, -3, 1) RING(14, -2, 2)  RING(15, -1, 3)
#undef RING
                const uint32_t co = WIN(4, 0), ce = WIN(3, 0);
#undef WIN
                const uint32_t se = arc_pair(re, ce, t_lo);    // pixels 0, 2
                const uint32_t so = arc_pair(ro, co, t_lo);    // pixels 1, 3
                *o = (so * 256u + se) & keep;
            }
        }
    }
#endif
    FAST_MARK(4);          // scoring
    __syncthreads();
    FAST_MARK(5);          // barrier after scoring

    // ---- phase B: one warp per cell
    const int warp = tid >> 5, lane = tid & 31;
    if (warp >= ncell) return;
    const int cj = j0 + warp;
    const int iniX = kMinBorder + cj * g.wCell;
    const int maxX = min(iniX + g.wCell + 6, g.maxBX);
    const int wd = (maxX - iniX) - 6;
    const long long cell = g.cell_base + (long long)ci * g.nCols + cj;
    int32_t* out_count = A.cell_count + (long long)blockIdx.y * A.cells_total + cell;
    uint32_t* out_stage = A.stage + (long long)blockIdx.y * A.stage_total + g.stage_base +
                          ((long long)ci * g.nCols + cj) * g.cell_cap;
    if (wd <= 0) return;
    // (the corner lists reuse the shifted strip copies, which nobody reads after phase A)
    uint16_t* clist = reinterpret_cast<uint16_t*>(smem + A.strip_bytes + warp * A.list_bytes);
    const int sc0 = SP + (iniX + 3 - xorg);    // score-map offset of the cell's detection pixel (0,0)

    // B1: row-major list (local index = ly<<6 | lx) of the scores that beat both in-row neighbours (a neighbour in
    //     another cell's columns counts as 0): one lane per detection row compares its row 4 pixels at a time (SWAR
    //     byte-wise >), keeps a bit per survivor, a warp prefix sum gives every row its offset, then each lane appends
    //     its row.  Only ~1/3 of the non-zero scores get this far.
    int ncorner = 0;
    const int nw = (wd + 3) >> 2;
    const uint32_t tail_mask = (wd & 3) ? ((1u << (8 * (wd & 3))) - 1u) : 0xffffffffu;
    for (int rbase = 0; rbase < hd; rbase += 32) {
        const int row = rbase + lane;
        uint32_t m_lo = 0, m_hi = 0;               // survivor bits of columns 0..31 / 32..63
        if (row < hd) {
            const int rb = sc0 + row * SP;
            uint32_t prev = 0, v = ld4(score, rb);
            if (nw == 1) v &= tail_mask;
            for (int wx = 0; wx < nw; wx++) {
                uint32_t next = 0;
                if (wx + 1 < nw) { next = ld4(score, rb + 4 * wx + 4); if (wx + 2 == nw) next &= tail_mask; }
                const uint32_t yl = __funnelshift_r(prev, v, 24), yr = __funnelshift_r(v, next, 8);   // left / right neighbours
                const uint32_t xl = v & 0x7f7f7f7fu;
                const uint32_t tl = (yl | 0x80808080u) - xl, tr = (yr | 0x80808080u) - xl;   // bit 7: low 7 bits of y >= those of v
                const uint32_t gel = (yl & ~v) | (~(yl ^ v) & tl), ger = (yr & ~v) | (~(yr ^ v) & tr);   // bit 7: neighbour >= v
                const uint32_t gt = ~(gel | ger) & 0x80808080u;                                    // bit 7: v > both
                const uint32_t nib = (((gt >> 7) * 0x00204081u) >> 21) & 0xfu;
                if (wx < 8) m_lo |= nib << (4 * wx); else m_hi |= nib << (4 * wx - 32);
                prev = v; v = next;
            }
        }
        const int cnt = __popc(m_lo) + __popc(m_hi);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int pos = ncorner + incl - cnt;
        while (m_lo) { clist[pos++] = (uint16_t)((row << 6) | (__ffs(m_lo) - 1)); m_lo &= m_lo - 1; }
        while (m_hi) { clist[pos++] = (uint16_t)((row << 6) | (__ffs(m_hi) + 31)); m_hi &= m_hi - 1; }
        ncorner += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();

    // B2 + B3: non-maximum suppression of the survivors against the rows above and below and ordered emission in ONE sweep at iniThFAST (the emitted
    //          corners are compacted in place at the front of the corner list); a cell without a single keypoint there
    //          is swept again at minThFAST -- nothing was overwritten in that case.  Rare on textured imagery.
    int cnt = 0;
    for (int pass = 0; pass < 2 && cnt == 0; pass++) {
        const int th = pass ? A.min_th : A.ini_th;
        for (int base = 0; base < ncorner; base += 32) {
            const int i = base + lane;
            bool emit = false; int idx = 0, s = 0;
            if (i < ncorner) {
                idx = clist[i];
                const int lx = idx & 63;
                const uint8_t* q = score + sc0 + (idx >> 6) * SP + lx;
                s = q[0];
                if (s >= th) {
                    const bool hasl = lx > 0, hasr = lx < wd - 1;        // neighbours in another cell's columns count as 0
                    const int l0 = hasl ? q[-SP - 1] : 0, l2 = hasl ? q[SP - 1] : 0;   // (B1 has compared the row neighbours)
                    const int r0 = hasr ? q[-SP + 1] : 0, r2 = hasr ? q[SP + 1] : 0;
                    emit = s > l0 && s > l2 && s > r0 && s > r2 && s > q[-SP] && s > q[SP];
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, emit);     // (also orders the reads above before the writes below)
            if (emit) {
                const int e = cnt + __popc(b & ((1u << lane) - 1));  // e <= i: the slot was read already
                out_stage[e] = (uint32_t)((idx & 63) + 3) | ((uint32_t)((idx >> 6) + 3) << 8) | ((uint32_t)s << 16);
                clist[e] = (uint16_t)idx;
            }
            cnt += __popc(b);
        }
        __syncwarp();
    }
    // B4: K3's count grid.  Every emitted key is counted in its depth-D cell of DivideNode's fixed grid and competes for
    //     that cell's best key (largest response, earliest emission), one atomic pair per distinct grid cell and sweep.
    {
        int32_t* hist = A.hist + (long long)blockIdx.y * A.hist_total + g.hist_base;
        unsigned long long* gbest = A.gbest + (long long)blockIdx.y * A.hist_total + g.hist_base;
        const int ox = cj * g.wCell + 3, oy = ci * g.hCell + 3;  // key coordinates are relative to (16,16)
        const unsigned long long cl = (unsigned long long)(ci * g.nCols + cj);
        for (int base = 0; base < cnt; base += 32) {
            const int e = base + lane;
            unsigned code = 0xffffffffu, v = 0;
            if (e < cnt) {
                const int idx = clist[e];
                const unsigned xl = __ldg(A.xlut + ox + (idx & 63));
                code = ((xl >> 8) << (2 * g.qt_depth)) + ((unsigned)__ldg(A.ylut + oy + (idx >> 6)) << g.qt_depth) + (xl & 0xff);
                v = ((unsigned)score[sc0 + (idx >> 6) * SP + (idx & 63)] << 12) | (unsigned)(0xfff - e);
            }
            const unsigned peers = __match_any_sync(0xffffffffu, code);
            const unsigned mv = __reduce_max_sync(peers, v);
            if (e < cnt && lane == __ffs(peers) - 1) {
                atomicAdd(hist + code, __popc(peers));
                atomicMax(gbest + code, ((unsigned long long)(mv >> 12) << 56) | (kBestOrderMask - ((cl << 12) | (0xfff - (mv & 0xfff)))));
            }
        }
    }
    if (lane == 0) *out_count = cnt;
    FAST_MARK(6);          // per-cell listing
}

}  // namespace

int fast_profile_read(unsigned long long* out16, int reset) {
#ifdef DSX_FAST_PROFILE
    std::vector<unsigned long long> h(512 * 16);
    if (cudaMemcpyFromSymbol(h.data(), g_fast_prof, sizeof(unsigned long long) * h.size()) != cudaSuccess) return DSX_ERR_CUDA;
    for (int k = 0; k < 16; k++) { out16[k] = 0; for (int s = 0; s < 512; s++) out16[k] += h[s * 16 + k]; }
    if (reset) {
        std::fill(h.begin(), h.end(), 0ull);
        if (cudaMemcpyToSymbol(g_fast_prof, h.data(), sizeof(unsigned long long) * h.size()) != cudaSuccess) return DSX_ERR_CUDA;
    }
    return DSX_OK;
#else
    (void)out16; (void)reset;
    set_error("built without DSX_FAST_PROFILE");
    return DSX_ERR_INVALID;
#endif
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        cudaGetLastError();
        return (EncodeTiledFn)p;
    }();
    return fn;
}
}  // namespace

// {pitch / 4 words, rows, n images} over the planes of one level; box = {SP / 4, box_rows, 1}
bool make_plane_map(CUtensorMap* m, const uint8_t* base, int pitch, int rows, long long img_stride, int n, int SP, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc || SP / 4 > 256 || box_rows > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)(pitch / 4), (cuuint64_t)rows, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)img_stride};
    const cuuint32_t box[3] = {(cuuint32_t)(SP / 4), (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_fast(dsx_ctx* ctx, const uint8_t* images, size_t step, size_t img_stride, int n) {
    StageTimer _t(ctx, 1);
    const ShapePlan& P = ctx->plan;
    DSX_CUDA(cudaMemsetAsync(ctx->ws.cell_count, 0, sizeof(int32_t) * (size_t)n * P.cells_total, ctx->stream));
    DSX_CUDA(cudaMemsetAsync(ctx->ws.hist, 0, sizeof(int32_t) * (size_t)n * P.hist_total, ctx->stream));
    DSX_CUDA(cudaMemsetAsync(ctx->ws.gbest, 0, sizeof(unsigned long long) * (size_t)n * P.hist_total, ctx->stream));
    // one strip pitch for all the levels of a launch (the widest cell decides), so that the kernel can have it as a constant
    auto level_sp = [&](int l) { return ((((kWarps * P.lv[l].wCell + 6 + 3 + 3) & ~3) + kPad + 8 + 12) + 15) & ~15; };
    int sp_launch = 0;
    auto fill = [&](int l, FastArgs& A) -> size_t {
        const LevelGeom& g = P.lv[l];
        A.g = g;
        A.img = (l == 0) ? images : ctx->ws.pyr + g.offset;
        A.img_stride = (l == 0) ? (long long)img_stride : P.pyr_bytes;
        A.pitch = (l == 0) ? (int)step : g.pitch;
        A.cell_count = ctx->ws.cell_count; A.stage = ctx->ws.stage;
        A.cells_total = P.cells_total; A.stage_total = P.stage_total;
        A.hist = ctx->ws.hist; A.gbest = ctx->ws.gbest; A.hist_total = P.hist_total;
        A.xlut = P.d_xlut + g.lut_x; A.ylut = P.d_ylut + g.lut_y;
        A.ini_th = std::min(std::max(ctx->p.ini_th_fast, 0), 255);
        A.min_th = std::min(std::max(ctx->p.min_th_fast, 0), 255);
        A.SP = sp_launch;
        A.strip_bytes = ((A.SP * (g.hCell + 6) + 16) + 15) & ~15;
        A.score_bytes = ((A.SP * (g.hCell + 2) + 16) + 15) & ~15;
        A.list_bytes = ((((g.wCell + 1) / 2) * g.hCell * 2) + 15) & ~15;    // in-row pre-suppression: <= ceil(w/2) entries per row
        // bulk-copy staging needs 16-byte aligned row segments: base, pitch and plane stride multiples of 16
        A.use_tma = (ctx->fast_tma && ((uintptr_t)A.img & 15) == 0 && (A.pitch & 15) == 0 && (A.img_stride & 15) == 0) ? 1 : 0;
        memset(&A.tmap, 0, sizeof(A.tmap));
        // DSX_FAST_TMA: 0 = 4-byte cp.async copies, 1 = one bulk copy per strip row, 2 (default) = one tensor-map box per strip
        if (A.use_tma && ctx->fast_tma >= 2 && make_plane_map(&A.tmap, A.img, A.pitch, g.rows, A.img_stride, n, A.SP, g.hCell + 6))
            A.use_tma = 2;
#if DSX_FAST_FP16
        return (size_t)5 * A.strip_bytes;                  // raw strip (later the score map) + two planes of 16-bit lanes
#else
        return (size_t)4 * A.strip_bytes + A.score_bytes;
#endif
    };
    // levels go out in groups of kFastMaxLevels per launch (one launch for the usual 6 levels)
    for (int l0 = 0; l0 < P.nlevels; l0 += kFastMaxLevels) {
        FastArgsAll AA;
        memset(&AA, 0, sizeof(AA));
        size_t smem = 0;
        int nl = 0, blocks = 0;
        sp_launch = 0;
        for (int l = l0; l < std::min(P.nlevels, l0 + kFastMaxLevels); l++)
            if (P.lv[l].n_cells) sp_launch = std::max(sp_launch, level_sp(l));
        for (int l = l0; l < std::min(P.nlevels, l0 + kFastMaxLevels); l++) {
            const LevelGeom& g = P.lv[l];
            if (g.n_cells == 0) continue;
            FastArgs& A = AA.lv[nl];
            const size_t sm = fill(l, A);
            if ((size_t)kWarps * A.list_bytes > (size_t)(DSX_FAST_FP16 ? 4 : 3) * A.strip_bytes) { set_error("FAST corner lists do not fit the strip copies"); return DSX_ERR_INVALID; }
            if (sm > 200 * 1024) { set_error("FAST cell too large for shared memory"); return DSX_ERR_INVALID; }
            smem = std::max(smem, sm);
            AA.first[nl] = blocks;
            blocks += g.nRows * ((g.nCols + kWarps - 1) / kWarps);
            nl++;
        }
        if (nl == 0) continue;
        AA.first[nl] = blocks;
        AA.nlevels = nl;
        dim3 grid(blocks, n);
        auto launch = [&](auto kernel) -> int {
            if (smem > 48 * 1024) DSX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kernel<<<grid, kWarps * 32, smem, ctx->stream>>>(AA);
            return DSX_OK;
        };
        int rc;
        switch (sp_launch) {          // the pitches of 30..33-pixel cells (every level of a usual survey) are compiled in
            case 288: rc = launch(fast_cells_kernel<288>); break;
            case 304: rc = launch(fast_cells_kernel<304>); break;
            default:  rc = launch(fast_cells_kernel<0>); break;
        }
        if (rc != DSX_OK) return rc;
        DSX_LAUNCH_CHECK();
    }
    return DSX_OK;
}

}  // namespace dsx
