// Host-side plan: everything ORBextractor's constructor and the per-level loop headers compute that
// does not depend on pixel values -- scale tables and quotas (ORBextractor.cpp:415-446), umax (:452-469),
// level sizes (:1119-1120), the FAST cell grid (:773-787), DistributeOctTree's root layout (:543-563) and
// the fixed-point INTER_LINEAR coefficient tables of cv::resize (OpenCV imgproc/resize.cpp, 11-bit
// coefficients).  All of it is float/double arithmetic evaluated on the host exactly as the reference does.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "dsx_internal.cuh"

namespace dsx {

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (v < (float)i); }

void init_tables(dsx_ctx* ctx) {
    const dsx_params& p = ctx->p;
    const int nl = p.nlevels;
    ctx->nlevels = nl;
    const double sf = (double)p.scale_factor;           // ORBextractor.h: `double scaleFactor` holds the float argument
    ctx->scale[0] = 1.0f; ctx->sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
        ctx->scale[i] = (float)(ctx->scale[i - 1] * sf);
        ctx->sigma2[i] = ctx->scale[i] * ctx->scale[i];
    }
    for (int i = 0; i < nl; i++) {
        ctx->inv_scale[i] = 1.0f / ctx->scale[i];
        ctx->inv_sigma2[i] = 1.0f / ctx->sigma2[i];
    }
    const float factor = (float)(1.0f / sf);
    float per_scale = p.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        ctx->quota[l] = cv_round_f(per_scale);
        sum += ctx->quota[l];
        per_scale *= factor;
    }
    ctx->quota[nl - 1] = std::max(p.nfeatures - sum, 0);
    // circular patch row ends
    int* umax = ctx->umax;
    const int vmax = cv_floor_f(kHalfPatch * sqrtf(2.f) / 2 + 1);
    const int vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// One axis of cv::resize INTER_LINEAR 8U: entry = src index | a1<<16 | inc<<31, a0 = 2048 - a1.
// `clamp_frac` reproduces the x axis (fraction zeroed at the borders); the y axis keeps its fraction and
// clips the two source rows instead (resizeGeneric_Invoker).
static void resize_axis_table(int n_dst, int n_src, bool is_x, uint32_t* tab) {
    const double scale = 1. / ((double)n_dst / n_src);
    for (int d = 0; d < n_dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor_f(f);
        f -= s;
        int s0, s1;
        if (is_x) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= n_src - 1) { f = 0; s = n_src - 1; }
            s0 = s; s1 = std::min(s + 1, n_src - 1);
        } else {
            s0 = std::min(std::max(s, 0), n_src - 1);
            s1 = std::min(std::max(s + 1, 0), n_src - 1);
        }
        const int a1 = cv_round_f(f * 2048.f);
        // a0 = cvRound((1-f)*2048) == 2048 - a1 for every float f in [0,1) (both products are exact)
        tab[d] = (uint32_t)s0 | ((uint32_t)a1 << 16) | ((uint32_t)(s1 - s0) << 31);
    }
}

void free_plan(dsx_ctx* ctx) {
    if (ctx->plan.d_tab) cudaFree(ctx->plan.d_tab);
    if (ctx->plan.d_xlut) cudaFree(ctx->plan.d_xlut);
    if (ctx->plan.d_ylut) cudaFree(ctx->plan.d_ylut);
    ctx->plan = ShapePlan();
}

static inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

int build_plan(dsx_ctx* ctx, int rows, int cols) {
    ShapePlan& P = ctx->plan;
    if (P.rows == rows && P.cols == cols && P.nlevels == ctx->nlevels) return DSX_OK;
    if (rows > kMaxDim || cols > kMaxDim) { set_error("image larger than 32760 in one dimension"); return DSX_ERR_INVALID; }
    DSX_CUDA(cudaStreamSynchronize(ctx->stream));
    free_plan(ctx);
    P.rows = rows; P.cols = cols; P.nlevels = ctx->nlevels;
    long long tab_total = 0, lut_x_total = 0, lut_y_total = 0;
    int key_base = 0;
    for (int l = 0; l < ctx->nlevels; l++) {
        LevelGeom& g = P.lv[l];
        std::memset(&g, 0, sizeof(g));
        const float s = ctx->inv_scale[l];
        g.cols = cv_round_f((float)cols * s);
        g.rows = cv_round_f((float)rows * s);
        if (g.rows < 2 * kMinBorder + 1 || g.cols < 2 * kMinBorder + 1) {
            // reference: (maxBorder-minBorder) <= 0 -> round(x/0) in DistributeOctTree; undefined
            set_error("image too small: a pyramid level is smaller than 33 pixels");
            P.rows = P.cols = 0;
            return DSX_ERR_INVALID;
        }
        g.pitch = (int)align_up(g.cols, 16);
        if (l > 0) { g.offset = P.pyr_bytes; P.pyr_bytes += align_up((long long)g.pitch * g.rows, 256); }
        g.maxBX = g.cols - kEdge + 3;
        g.maxBY = g.rows - kEdge + 3;
        const float width = (float)(g.maxBX - kMinBorder), height = (float)(g.maxBY - kMinBorder);
        g.nCols = (int)(width / kCellW);
        g.nRows = (int)(height / kCellW);
        g.wCell = g.nCols > 0 ? (int)ceilf(width / g.nCols) : 0;
        g.hCell = g.nRows > 0 ? (int)ceilf(height / g.nRows) : 0;
        if (g.nCols == 0 || g.nRows == 0) g.nCols = g.nRows = 0;  // the reference's cell loops are empty
        g.n_cells = g.nCols * g.nRows;
        g.cell_cap = std::max(1, ((g.wCell + 1) / 2) * ((g.hCell + 1) / 2));
        g.cell_base = P.cells_total; P.cells_total += g.n_cells;
        g.stage_base = P.stage_total; P.stage_total += (long long)g.n_cells * g.cell_cap;
        const long long worst = (long long)g.n_cells * g.cell_cap;
        g.cand_cap = (int)std::min<long long>(worst, std::max<long long>(4096, (long long)g.rows * g.cols / 8));
        g.cand_cap = std::max(g.cand_cap, 1);
        g.cand_base = P.cand_total; P.cand_total += align_up(g.cand_cap, 4);
        g.quota = ctx->quota[l];
        const int w = g.maxBX - kMinBorder, h = g.maxBY - kMinBorder;
        g.nIni = (int)roundf((float)w / h);
        if (g.nIni < 1) g.nIni = 1;  // B1
        g.hX = (float)w / g.nIni;
        g.node_cap = std::max(g.quota + 2, 4 * g.nIni);
        g.key_base = key_base; key_base += g.node_cap;
        g.scale = ctx->scale[l];
        g.kp_size = (float)(int)(31 * ctx->scale[l]);
        P.max_roi_w = std::max(P.max_roi_w, g.wCell + 6);
        P.max_roi_h = std::max(P.max_roi_h, g.hCell + 6);
        if (l > 0) { P.xtab_off[l] = tab_total; tab_total += g.cols; P.ytab_off[l] = tab_total; tab_total += g.rows; }
        // count-grid depth: 6 covers 4096 cells per root (final nodes sit near depth log4(quota) ~ 4.5 for evenly spread
        // keys) and keeps the quadtree CTA under half an SM's shared memory
        if (g.nIni > 255) { set_error("aspect ratio too extreme: more than 255 quadtree root nodes"); P.rows = P.cols = 0; return DSX_ERR_INVALID; }
        int D = 6;
        g.qt_global = 0;
        while (D > 0 && quadtree_smem_bytes(g, D) > 110 * 1024) D--;
        if (quadtree_smem_bytes(g, D) > 110 * 1024 || (D < 5 && g.node_cap > 1024)) {
            // large quota (nfeatures beyond ~8000): node arrays in global scratch, a deeper grid in shared memory
            g.qt_global = 1;
            D = 7;
            while (D > 0 && quadtree_smem_bytes(g, D) > 200 * 1024) D--;
        }
        if (quadtree_smem_bytes(g, D) > 220 * 1024 || g.node_cap > 65535) {
            set_error("quadtree: root count / quota exceed the kernel's limits (<= 65533 keys per level)");
            P.rows = P.cols = 0;
            return DSX_ERR_INVALID;
        }
        P.qt_scratch_stride = std::max(P.qt_scratch_stride, quadtree_scratch_bytes(g));
        g.qt_depth = D;
        g.hist_base = P.hist_total; P.hist_total += (long long)g.nIni << (2 * D);
        g.lut_x = lut_x_total; lut_x_total += w;
        g.lut_y = lut_y_total; lut_y_total += h;
    }
    P.keys_total = key_base;
    {   // DivideNode halves a box at ceil((hi-lo)/2) (ORBextractor.cpp:483-484): D halvings of the root box give every
        // key column / row its depth-D grid cell; roots per :543-563, root of a key per :569
        std::vector<uint16_t> xl((size_t)std::max<long long>(lut_x_total, 1));
        std::vector<uint8_t> yl((size_t)std::max<long long>(lut_y_total, 1));
        for (int l = 0; l < ctx->nlevels; l++) {
            const LevelGeom& g = P.lv[l];
            const int w = g.maxBX - kMinBorder, h = g.maxBY - kMinBorder, D = g.qt_depth;
            for (int x = 0; x < w; x++) {
                const int r = (int)((float)x / g.hX);
                int lo = (int)(g.hX * (float)r), hi = (int)(g.hX * (float)(r + 1)), c = 0;
                for (int d = 0; d < D; d++) {
                    const int mid = lo + ((hi - lo + 1) >> 1);
                    if (x < mid) { hi = mid; c = c << 1; } else { lo = mid; c = (c << 1) | 1; }
                }
                xl[g.lut_x + x] = (uint16_t)((r << 8) | c);
            }
            for (int y = 0; y < h; y++) {
                int lo = 0, hi = h, c = 0;
                for (int d = 0; d < D; d++) {
                    const int mid = lo + ((hi - lo + 1) >> 1);
                    if (y < mid) { hi = mid; c = c << 1; } else { lo = mid; c = (c << 1) | 1; }
                }
                yl[g.lut_y + y] = (uint8_t)c;
            }
        }
        DSX_CUDA(cudaMalloc(&P.d_xlut, xl.size() * sizeof(uint16_t)));
        DSX_CUDA(cudaMalloc(&P.d_ylut, yl.size()));
        DSX_CUDA(cudaMemcpy(P.d_xlut, xl.data(), xl.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        DSX_CUDA(cudaMemcpy(P.d_ylut, yl.data(), yl.size(), cudaMemcpyHostToDevice));
    }
    std::vector<uint32_t> tab((size_t)std::max<long long>(tab_total, 1));
    for (int l = 1; l < ctx->nlevels; l++) {
        resize_axis_table(P.lv[l].cols, P.lv[l - 1].cols, true, tab.data() + P.xtab_off[l]);
        resize_axis_table(P.lv[l].rows, P.lv[l - 1].rows, false, tab.data() + P.ytab_off[l]);
        // K1's TMA-staged form picks the four tap pairs of an aligned group of 4 destination columns out of one 8-byte
        // window: their left taps must lie within 4 source bytes of the group's first one
        const uint32_t* xt = tab.data() + P.xtab_off[l];
        bool ok = true;
        for (int x = 0; x < P.lv[l].cols && ok; x += 4) {
            const int xe = std::min(x + 3, P.lv[l].cols - 1);
            ok = (int)(xt[xe] & 0xffff) - (int)(xt[x] & 0xffff) <= 4;
        }
        P.xspan4[l] = ok;
    }
    DSX_CUDA(cudaMalloc(&P.d_tab, tab.size() * sizeof(uint32_t)));
    DSX_CUDA(cudaMemcpyAsync(P.d_tab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    DSX_CUDA(cudaStreamSynchronize(ctx->stream));
    // K5 / K6 form.  The per-keypoint form reads one 49x49 window per keypoint (2401 B) and blurs 37x37 of it; the dense
    // form blurs every level once (5.8 B per pixel of the 2.9 RC pyramid pixels) and gathers.  The dense form is chosen
    // where it moves fewer bytes, N * 2401 > 5.8 * 2.906 * RC -- at the survey shapes that needs more than the 50 k
    // keypoints of BASELINE config 5 (8000 x 2000: N > 112 k), and the measured stage times agree (profiles/r02:
    // 14.4 ms against 23.7 ms at 20 k keypoints), so in practice it serves small, very dense images.
    // DSX_DESCRIBE_DENSE=0/1 forces one form (the tests run both).
    {
        long long o = 0;
        for (int l = 0; l < P.nlevels; l++) {
            P.blur_pitch[l] = (P.lv[l].cols + 15) & ~15;
            P.blur_off[l] = o;
            o += align_up((long long)P.blur_pitch[l] * P.lv[l].rows, 256);
        }
        P.blur_bytes = o;
        P.dense_describe = ((double)ctx->p.nfeatures * 2401.0 > 5.8 * 2.906 * (double)rows * cols) ? 1 : 0;
        if (const char* e = getenv("DSX_DESCRIBE_DENSE")) P.dense_describe = atoi(e) ? 1 : 0;
    }
    // workspace depends on the shape: drop it
    ctx->ws.batch = 0;
    return DSX_OK;
}

template <typename T>
static int re_alloc(T*& p, size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return DSX_ERR_NOMEM; }
    return DSX_OK;
}

int ensure_workspace(dsx_ctx* ctx, int batch) {
    Workspace& W = ctx->ws;
    const ShapePlan& P = ctx->plan;
    if (W.batch >= batch) return DSX_OK;
    DSX_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t B = (size_t)batch;
    DSX_TRY(re_alloc(W.pyr, B * (size_t)P.pyr_bytes));
    DSX_TRY(re_alloc(W.cell_count, B * (size_t)P.cells_total));
    DSX_TRY(re_alloc(W.stage, B * (size_t)P.stage_total));
    DSX_TRY(re_alloc(W.cand_xy, B * (size_t)P.cand_total));
    DSX_TRY(re_alloc(W.cand_resp, B * (size_t)P.cand_total));
    DSX_TRY(re_alloc(W.cand_node, B * (size_t)P.cand_total));
    DSX_TRY(re_alloc(W.cand_count, B * DSX_MAX_LEVELS));
    DSX_TRY(re_alloc(W.key_xy, B * (size_t)P.keys_total));
    DSX_TRY(re_alloc(W.key_resp, B * (size_t)P.keys_total));
    DSX_TRY(re_alloc(W.key_count, B * DSX_MAX_LEVELS));
    DSX_TRY(re_alloc(W.hist, B * (size_t)P.hist_total));
    DSX_TRY(re_alloc(W.gbest, B * (size_t)P.hist_total));
    DSX_TRY(re_alloc(W.deep, B * DSX_MAX_LEVELS));
    if (P.qt_scratch_stride) {
        uint8_t* sc = (uint8_t*)W.node_scratch;
        DSX_TRY(re_alloc(sc, B * (size_t)P.nlevels * P.qt_scratch_stride));
        W.node_scratch = sc;
    }
    if (P.dense_describe) DSX_TRY(re_alloc(W.blur, B * (size_t)P.blur_bytes));
    DSX_TRY(re_alloc(W.tmp_kps, B * (size_t)ctx->cap));
    DSX_TRY(re_alloc(W.tmp_desc, B * (size_t)ctx->cap * 32));
    DSX_TRY(re_alloc(W.tmp_count, B));
    if (!W.err_flag) {
        DSX_TRY(re_alloc(W.err_flag, 1));
        DSX_CUDA(cudaMemsetAsync(W.err_flag, 0, sizeof(int32_t), ctx->stream));
    }
    W.batch = batch;
    return DSX_OK;
}

}  // namespace dsx
