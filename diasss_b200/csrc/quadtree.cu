// K3 -- keypoint distribution.  Replaces ORBextractor::DistributeOctTree (ORBextractor.cpp:539-763) and
// ExtractorNode::DivideNode (:481-537), plus the candidate append order of the cell loop (:818-826).
//
// The reference keeps a std::list of nodes and pushes children to its front.  This kernel keeps the same
// list as an ARRAY whose index is the list position, and derives every new position arithmetically:
//
//   normal pass (:606-665)  every node with >1 keys is split, visiting the list front to back; non-empty
//       children are push_front-ed in the order n1..n4.  New list = [children of the LAST visited node
//       (n4,n3,n2,n1), ..., children of the FIRST visited node] ++ [single-key nodes in their old order].
//   final phase (:673-738)  nodes with >1 keys are visited by (key count descending, most recently created
//       first [B2] == smaller list position first) until the list reaches N nodes; same push_front rule;
//       unvisited nodes keep their relative order behind the new children.
//   stop when size >= N or a step adds no node (:669, :734); switch to the final phase when
//       size + 3*nToExpand > N (:673).
//   result (:742-760)  per node the key with the largest response, the earliest key among equals
//       (candidate order = cell-row-major, row-major inside a cell), in list order.
//
// Child key counts.  DivideNode's geometry (ceil-halving of the parent box, :483-484) does not depend on the keys,
// so every node down to depth D is a cell of a fixed, non-uniform 2^d x 2^d grid per root.  K2 histograms every key
// it emits into that depth-D grid (global memory, warp-aggregated atomics), so the distribution never has to walk the
// keys:
//   quadtree_kernel<false>  one CTA per (image, level): loads the grid, a 2x2 reduction builds the count pyramid, and
//       the children counts of any node above depth D are table look-ups; the whole list evolution runs on
//       <= quota+2 node records in shared memory (block scans; a bitonic sort in the final phase).  K2 also keeps,
//       per grid cell, the best key (64-bit atomicMax of response << 56 | earliest emission order, warp-aggregated);
//       a final node above depth D is a block of grid cells, so its winner is a max over that block -- the keys
//       themselves are never revisited.
// Only if a node AT depth D must be split (strongly clustered keys) does an (image, level) fall back to the general
// form, quadtree_kernel<true>: it builds the reference-ordered candidate list from K2's per-cell slots, each key
// carries (list position << 2 | quadrant), and one sweep per step applies the previous step's position remap and
// counts children with atomics.  That kernel is always launched and exits at once for (image, level)s that did not
// raise the `deep` flag.
#include "dsx_internal.cuh"

namespace dsx {

namespace {

constexpr int kThreads = 1024;

struct QtArgs {
    LevelGeom lv[DSX_MAX_LEVELS];
    int nlevels;
    int32_t* hist; unsigned long long* gbest; int32_t* deep;
    const uint16_t* xlut; const uint8_t* ylut;
    long long hist_total;
    uint8_t* node_scratch; long long node_scratch_stride;   // node arrays of (image, level)s whose quota does not fit shared memory
    int32_t* cell_count; uint32_t* stage;
    uint32_t* cand_xy; uint8_t* cand_resp; uint32_t* cand_node; int32_t* cand_count;
    uint32_t* key_xy; uint8_t* key_resp; int32_t* key_count;
    long long cells_total, stage_total, cand_total;
    int keys_total;
    int32_t* err_flag;
};

__device__ __forceinline__ int block_sum_scan(int v, int* wsum, int* total) {
    // exclusive scan of one value per thread over the block; *total = block sum
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = wsum[lane], i2 = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, i2, o); if (lane >= o) i2 += t; }
        wsum[lane] = i2 - s;
        if (lane == 31) wsum[32] = i2;
    }
    __syncthreads();
    *total = wsum[32];
    return wsum[w] + incl - v;
}

// in-place exclusive scan of a[0..n) (shared or global); returns the total to every thread
__device__ int block_excl_scan(int* a, int n, int* wsum) {
    const int chunk = (n + kThreads - 1) / kThreads;
    const int b = min(threadIdx.x * chunk, n), e = min(b + chunk, n);
    int s = 0;
    for (int i = b; i < e; i++) s += a[i];
    int total;
    int run = block_sum_scan(s, wsum, &total);
    for (int i = b; i < e; i++) { int t = a[i]; a[i] = run; run += t; }
    __syncthreads();
    return total;
}

__device__ void bitonic_sort_desc(unsigned long long* k, int n2) {
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (n2 >> 1); i += kThreads) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = k[lo], b = k[hi];
                if ((a < b) == desc) { k[lo] = b; k[hi] = a; }
            }
        }
    __syncthreads();
}

__device__ __forceinline__ int pyr_base(int d) { return ((1 << (2 * d)) - 1) / 3; }   // sum_{e<d} 4^e

// bytes of the node-sized arrays of one (image, level): sort keys, 2x(box, count, meta), children, 5 scan arrays, remap
__host__ __device__ inline size_t quadtree_node_bytes(int NC) {
    int n2 = 1; while (n2 < NC) n2 <<= 1;
    return ((size_t)n2 * 8 + (size_t)NC * (8 + 8 + 4 + 4 + 4 + 4 + 16 + 4 * 5 + 8) + 255) & ~(size_t)255;
}

template <bool GENERAL>
__global__ void __launch_bounds__(kThreads, 1) quadtree_kernel(const QtArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int level = blockIdx.x, img = blockIdx.y;
    if (GENERAL && !A.deep[img * DSX_MAX_LEVELS + level]) return;
    const LevelGeom& g = A.lv[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NC = g.node_cap;
    const int D = g.qt_depth;
    int n2 = 1; while (n2 < NC) n2 <<= 1;
    const int h = g.maxBY - kMinBorder;
    const int pyr_per_root = pyr_base(D + 1);
    const int cells_per_root = 1 << (2 * D);

    // ---- carve-up (sizes mirrored in quadtree_node_bytes / quadtree_fixed_bytes): the node-sized arrays live in shared
    //      memory, or -- for quotas beyond ~2500 keys per level -- in a global scratch block of this (image, level); the
    //      count pyramid and the cell table always sit in shared memory
    uint8_t* nbase = g.qt_global ? A.node_scratch + ((long long)img * A.nlevels + level) * A.node_scratch_stride : smem;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(nbase);  // [n2] final-phase sort keys
    short4* nodeA = reinterpret_cast<short4*>(skey + n2);                      // [NC] x0,y0,x1,y1
    short4* nodeB = nodeA + NC;
    int* cntA = reinterpret_cast<int*>(nodeB + NC);
    int* cntB = cntA + NC;
    unsigned* metaA = reinterpret_cast<unsigned*>(cntB + NC);                  // [NC] cx | cy<<10 | depth<<20 | root<<24
    unsigned* metaB = metaA + NC;
    int* child = reinterpret_cast<int*>(metaB + NC);   // [NC][4] children key counts; reused as best[NC] in the last sweep
    int* nch = child + 4 * NC;         // non-empty children per node
    int* gpos = nch + NC;              // first list position of a split node's children
    int* kscan = gpos + NC;            // scan scratch
    int* splitf = kscan + NC;          // 1 = node is split in this step
    int* ridx = splitf + NC;           // final phase: visiting rank -> list position
    uint16_t* remap = reinterpret_cast<uint16_t*>(ridx + NC);                  // [NC][4] old (pos,quadrant) -> new pos
    int* wsum = reinterpret_cast<int*>(g.qt_global ? smem : nbase + quadtree_node_bytes(NC));   // 33 ints
    int* pyr = wsum + 36;                                                      // [nIni][pyr_per_root] count pyramid
    uint16_t* cellnode = reinterpret_cast<uint16_t*>(pyr + g.nIni * pyr_per_root);   // [nIni][4^D] depth-D cell -> list position
    __shared__ int s_np, s_nexp, s_deep;

    int32_t* cell_cnt = A.cell_count + (long long)img * A.cells_total + g.cell_base;
    const uint32_t* stage = A.stage + (long long)img * A.stage_total + g.stage_base;
    uint32_t* cxy = A.cand_xy + (long long)img * A.cand_total + g.cand_base;
    uint8_t* cresp = A.cand_resp + (long long)img * A.cand_total + g.cand_base;
    uint32_t* cnode = A.cand_node + (long long)img * A.cand_total + g.cand_base;
    const int32_t* hist = A.hist + (long long)img * A.hist_total + g.hist_base;
    const uint16_t* xlut = A.xlut + g.lut_x;                                   // [w] root<<8 | depth-D column
    const uint8_t* ylut = A.ylut + g.lut_y;                                    // [h] depth-D row

    const int N = g.quota;
    const float hX = g.hX;
    // ---- depth-D level of the count pyramid = K2's histogram
    const int baseD = pyr_base(D);
    for (int i = tid; i < g.nIni * cells_per_root; i += kThreads) {
        const int r = i >> (2 * D);
        pyr[r * pyr_per_root + baseD + (i & (cells_per_root - 1))] = hist[i];
    }
    int n = 0;
    if (GENERAL) {
        // ---- reference-ordered candidate list (cell-row-major, row-major inside a cell; ORBextractor.cpp:818-826)
        n = block_excl_scan(cell_cnt, g.n_cells, wsum);   // cell_cnt[c] becomes the offset of cell c (syncs inside)
        const int n_all = n;
        if (n > g.cand_cap) {
            if (tid == 0) atomicExch(A.err_flag, DSX_ERR_CAPACITY);
            n = g.cand_cap;
        }
        for (int c = warp; c < g.n_cells; c += kThreads / 32) {
            const int off = cell_cnt[c];
            const int end = min((c + 1 < g.n_cells) ? cell_cnt[c + 1] : n_all, n);
            const int ci = c / g.nCols, cj = c - ci * g.nCols;
            const uint32_t* src = stage + (long long)c * g.cell_cap;
            const int ox = cj * g.wCell, oy = ci * g.hCell;
            for (int e = lane; off + e < end; e += 32) {
                const uint32_t p = src[e];
                cxy[off + e] = ((p & 0xff) + ox) | ((((p >> 8) & 0xff) + oy) << 16);
                cresp[off + e] = (uint8_t)(p >> 16);
            }
        }
        if (tid == 0) A.cand_count[img * DSX_MAX_LEVELS + level] = n;
    }
    __syncthreads();
    for (int d = D - 1; d >= 0; d--) {                     // 2x2 reduction: counts of every node of the fixed grids
        const int per = 1 << (2 * d);
        for (int i = tid; i < g.nIni * per; i += kThreads) {
            const int r = i / per, e = i - r * per, cy = e >> d, cx = e & ((1 << d) - 1);
            const int* up = pyr + r * pyr_per_root + pyr_base(d + 1);
            const int o = ((2 * cy) << (d + 1)) + 2 * cx;
            pyr[r * pyr_per_root + pyr_base(d) + e] = up[o] + up[o + 1] + up[o + (2 << d)] + up[o + (2 << d) + 1];
        }
        __syncthreads();
    }
    // initial list: non-empty roots in index order (:552-585)
    for (int i = tid; i < g.nIni; i += kThreads) kscan[i] = pyr[i * pyr_per_root] > 0;
    __syncthreads();
    int m = block_excl_scan(kscan, g.nIni, wsum);
    for (int i = tid; i < g.nIni; i += kThreads) {
        const int c = pyr[i * pyr_per_root];
        if (c > 0) {
            const int p = kscan[i];
            nodeB[p] = make_short4((short)(int)__fmul_rn(hX, (float)i), 0, (short)(int)__fmul_rn(hX, (float)(i + 1)), (short)h);
            cntB[p] = c;
            metaB[p] = (unsigned)i << 24;
        }
    }
    __syncthreads();
    short4* cur = nodeB; short4* nxt = nodeA;
    int* ccnt = cntB; int* ncnt = cntA;
    unsigned* cmeta = metaB; unsigned* nmeta = metaA;

    bool final_phase = false, finished = (m == 0), keyed = false;   // keyed: general form (keys carry their node)
    while (!finished) {
        if (tid == 0) { s_nexp = 0; s_np = 0x7fffffff; s_deep = 0; }
        __syncthreads();
        if (!keyed) {
            // children counts from the pyramid; does any splittable node sit at depth D?
            for (int i = tid; i < m; i += kThreads) {
                if (ccnt[i] > 1) {
                    const unsigned mt = cmeta[i];
                    const int d = (mt >> 20) & 15, cx = mt & 0x3ff, cy = (mt >> 10) & 0x3ff, r = mt >> 24;
                    if (d >= D) { s_deep = 1; }
                    else {
                        const int* up = pyr + r * pyr_per_root + pyr_base(d + 1);
                        const int o = ((2 * cy) << (d + 1)) + 2 * cx;
                        child[4 * i] = up[o]; child[4 * i + 1] = up[o + 1];
                        child[4 * i + 2] = up[o + (2 << d)]; child[4 * i + 3] = up[o + (2 << d) + 1];
                    }
                }
            }
            __syncthreads();
            if (!GENERAL && s_deep) {
                // a node of the depth-D grid must be split: this (image, level) is redone by quadtree_kernel<true>
                if (tid == 0) A.deep[img * DSX_MAX_LEVELS + level] = 1;
                return;
            }
            if (GENERAL && s_deep) {
                // ---- switch to the general form: give every key its current list position
                for (int i = warp; i < m; i += kThreads / 32) {
                    const unsigned mt = cmeta[i];
                    const int d = (mt >> 20) & 15, cx = mt & 0x3ff, cy = (mt >> 10) & 0x3ff, r = mt >> 24;
                    const int span = 1 << (D - d);
                    uint16_t* dst = cellnode + r * cells_per_root;
                    for (int e = lane; e < span * span; e += 32)
                        dst[(((cy << (D - d)) + e / span) << D) + (cx << (D - d)) + (e & (span - 1))] = (uint16_t)i;
                }
                for (int i = tid; i < 4 * m; i += kThreads) remap[i] = (uint16_t)(i >> 2);
                __syncthreads();
                for (int k = tid; k < n; k += kThreads) {
                    const uint32_t xy = cxy[k];
                    const int xl = xlut[xy & 0xffff];
                    cnode[k] = (unsigned)cellnode[(xl >> 8) * cells_per_root + (ylut[xy >> 16] << D) + (xl & 0xff)] << 2;
                }
                keyed = true;
                __syncthreads();
            }
        }
        if (GENERAL && keyed) {
            // ---- key sweep: apply the previous remap, count the children of every splittable node
            for (int i = tid; i < 4 * m; i += kThreads) child[i] = 0;
            __syncthreads();
            for (int base = 0; base < n; base += kThreads) {
                const int k = base + tid;
                unsigned code = 0xffffffffu;
                if (k < n) {
                    const int idx = remap[cnode[k]];
                    unsigned packed = (unsigned)idx << 2;
                    if (ccnt[idx] > 1) {
                        const uint32_t xy = cxy[k];
                        const int x = xy & 0xffff, y = xy >> 16;
                        const short4 b = cur[idx];
                        const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);   // ceil halves (:483-484)
                        packed |= (x < mx ? 0 : 1) + (y < my ? 0 : 2);                                    // n1..n4 (:512-526)
                        code = packed;
                    }
                    cnode[k] = packed;
                }
                const unsigned peers = __match_any_sync(0xffffffffu, code);
                if (code != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&child[code], __popc(peers));
            }
            __syncthreads();
        }

        // ---- which nodes are split, and where do their children go
        for (int i = tid; i < m; i += kThreads) {
            int c = 0;
            if (ccnt[i] > 1) c = (child[4 * i] > 0) + (child[4 * i + 1] > 0) + (child[4 * i + 2] > 0) + (child[4 * i + 3] > 0);
            nch[i] = c;
            splitf[i] = (!final_phase && ccnt[i] > 1) ? 1 : 0;
            kscan[i] = c;
        }
        __syncthreads();
        int total_ch;
        if (!final_phase) {
            total_ch = block_excl_scan(kscan, m, wsum);            // kscan[i] = children of nodes in front of i
            for (int i = tid; i < m; i += kThreads) gpos[i] = total_ch - kscan[i] - nch[i];
            __syncthreads();
        } else {
            for (int i = tid; i < n2; i += kThreads)
                skey[i] = (i < m && ccnt[i] > 1) ? (((unsigned long long)(unsigned)ccnt[i] << 32) | (0xffffffffu - (unsigned)i)) : 0ull;
            bitonic_sort_desc(skey, n2);
            for (int r = tid; r < m; r += kThreads) {
                const unsigned long long key = skey[r];
                const int idx = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffu)) : -1;
                ridx[r] = idx;
                kscan[r] = idx >= 0 ? nch[idx] - 1 : 0;
            }
            __syncthreads();
            block_excl_scan(kscan, m, wsum);                       // kscan[r] = list growth before visiting rank r
            for (int r = tid; r < m; r += kThreads) {
                const int idx = ridx[r];
                if (idx < 0) atomicMin(&s_np, r);                                          // no more splittable nodes
                else if (m + kscan[r] + nch[idx] - 1 >= N) atomicMin(&s_np, r + 1);        // break after this one (:730)
            }
            __syncthreads();
            const int np = min(s_np, m);
            for (int r = tid; r < m; r += kThreads) kscan[r] = (r < np) ? nch[ridx[r]] : 0;
            __syncthreads();
            total_ch = block_excl_scan(kscan, m, wsum);            // kscan[r] = children of ranks visited before r
            for (int r = tid; r < np; r += kThreads) {
                const int idx = ridx[r];
                splitf[idx] = 1;
                gpos[idx] = total_ch - kscan[r] - nch[idx];        // later visited = further to the front
            }
            __syncthreads();
        }
        for (int i = tid; i < m; i += kThreads) kscan[i] = splitf[i] ? 0 : 1;
        __syncthreads();
        const int n_keep = block_excl_scan(kscan, m, wsum);
        const int m_new = total_ch + n_keep;

        // ---- build the new list (and, in the general form, the remap table)
        int my_exp = 0;
        for (int i = tid; i < m; i += kThreads) {
            if (splitf[i]) {
                const short4 b = cur[i];
                const unsigned mt = cmeta[i];
                const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
                int pos = gpos[i];
#pragma unroll
                for (int q = 3; q >= 0; q--) {                     // n4 ends up foremost
                    const int c = child[4 * i + q];
                    if (c > 0) {
                        nxt[pos] = make_short4((short)((q & 1) ? mx : b.x), (short)((q & 2) ? my : b.y),
                                               (short)((q & 1) ? b.z : mx), (short)((q & 2) ? b.w : my));
                        ncnt[pos] = c;
                        // depth saturates at 15; cx/cy are only read while depth < D <= 6 (wrap-around beyond is harmless)
                        nmeta[pos] = (mt & 0xff000000u) | (min(((mt >> 20) & 15) + 1, 15u) << 20) |
                                     (((((mt >> 10) & 0x3ff) * 2 + (q >> 1)) & 0x3ff) << 10) | (((mt & 0x3ff) * 2 + (q & 1)) & 0x3ff);
                        remap[4 * i + q] = (uint16_t)pos;
                        my_exp += (c > 1);
                        pos++;
                    }
                }
            } else {
                const int pos = total_ch + kscan[i];
                nxt[pos] = cur[i];
                ncnt[pos] = ccnt[i];
                nmeta[pos] = cmeta[i];
#pragma unroll
                for (int q = 0; q < 4; q++) remap[4 * i + q] = (uint16_t)pos;
            }
        }
        if (my_exp) atomicAdd(&s_nexp, my_exp);
        __syncthreads();
        const int nexp = s_nexp;
        if (m_new >= N || m_new == m) finished = true;                     // :669 / :734
        else if (!final_phase && m_new + 3 * nexp > N) final_phase = true; // :673
        m = m_new;
        { short4* t = cur; cur = nxt; nxt = t; int* u = ccnt; ccnt = ncnt; ncnt = u; unsigned* v = cmeta; cmeta = nmeta; nmeta = v; }
        __syncthreads();
    }

    if (!GENERAL) {
        // ---- per final node the key with the largest response, the earliest emitted among equals (:742-760): a node above
        //      depth D is a block of depth-D grid cells whose best keys K2 has already determined
        const unsigned long long* gbest = A.gbest + (long long)img * A.hist_total + g.hist_base;
        uint32_t* kxy = A.key_xy + (long long)img * A.keys_total + g.key_base;
        uint8_t* kresp = A.key_resp + (long long)img * A.keys_total + g.key_base;
        for (int i = warp; i < m; i += kThreads / 32) {
            const unsigned mt = cmeta[i];
            const int d = (mt >> 20) & 15, cx = mt & 0x3ff, cy = (mt >> 10) & 0x3ff, r = mt >> 24;
            const int span = 1 << (D - d);
            const unsigned long long* src = gbest + r * cells_per_root + ((cy << (D - d)) << D) + (cx << (D - d));
            unsigned long long v = 0ull;
            for (int e = lane; e < span * span; e += 32) v = max(v, src[((e >> (D - d)) << D) + (e & (span - 1))]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) {
                const unsigned long long order = kBestOrderMask - (v & kBestOrderMask);
                const int cl = (int)(order >> 12), e = (int)(order & 0xfff);
                const int ci = cl / g.nCols, cj = cl - ci * g.nCols;
                const uint32_t p = stage[(long long)cl * g.cell_cap + e];
                kxy[i] = (uint32_t)((p & 0xff) + cj * g.wCell + kMinBorder) | ((uint32_t)(((p >> 8) & 0xff) + ci * g.hCell + kMinBorder) << 16);   // :843-844
                kresp[i] = (uint8_t)(v >> 56);
            }
        }
        if (tid == 0) A.key_count[img * DSX_MAX_LEVELS + level] = m;
        return;
    }
    // ---- general form, last sweep: per node the largest response, earliest candidate among equals (:742-760)
    unsigned* best = reinterpret_cast<unsigned*>(child);
    for (int i = tid; i < m; i += kThreads) best[i] = 0;
    if (!keyed) {
        for (int i = warp; i < m; i += kThreads / 32) {            // depth-D cell -> final list position
            const unsigned mt = cmeta[i];
            const int d = (mt >> 20) & 15, cx = mt & 0x3ff, cy = (mt >> 10) & 0x3ff, r = mt >> 24;
            const int span = 1 << (D - d);
            uint16_t* dst = cellnode + r * cells_per_root;
            for (int e = lane; e < span * span; e += 32)
                dst[(((cy << (D - d)) + e / span) << D) + (cx << (D - d)) + (e & (span - 1))] = (uint16_t)i;
        }
    }
    __syncthreads();
    for (int k = tid; k < n; k += kThreads) {
        int idx;
        if (keyed) idx = remap[cnode[k]];
        else {
            const uint32_t xy = cxy[k];
            const int xl = xlut[xy & 0xffff];
            idx = cellnode[(xl >> 8) * cells_per_root + (ylut[xy >> 16] << D) + (xl & 0xff)];
        }
        atomicMax(&best[idx], ((unsigned)cresp[k] << 24) | (0xffffffu - (unsigned)k));
    }
    __syncthreads();
    uint32_t* kxy = A.key_xy + (long long)img * A.keys_total + g.key_base;
    uint8_t* kresp = A.key_resp + (long long)img * A.keys_total + g.key_base;
    for (int i = tid; i < m; i += kThreads) {
        const unsigned v = best[i];
        const uint32_t xy = cxy[0xffffffu - (v & 0xffffffu)];
        kxy[i] = ((xy & 0xffff) + kMinBorder) | (((xy >> 16) + kMinBorder) << 16);    // :843-844
        kresp[i] = (uint8_t)(v >> 24);
    }
    if (tid == 0) A.key_count[img * DSX_MAX_LEVELS + level] = m;
}

}  // namespace

// shared memory that does not scale with the quota: scan scratch, count pyramid, depth-D cell table
static size_t quadtree_fixed_bytes(const LevelGeom& g, int D) {
    const size_t pyr = (size_t)g.nIni * (((size_t)1 << (2 * (D + 1))) - 1) / 3;
    const size_t cells = (size_t)g.nIni << (2 * D);
    return 36 * 4 + pyr * 4 + cells * 2 + 64;
}

size_t quadtree_smem_bytes(const LevelGeom& g, int D) {
    return (g.qt_global ? 0 : quadtree_node_bytes(g.node_cap)) + quadtree_fixed_bytes(g, D);
}

size_t quadtree_scratch_bytes(const LevelGeom& g) { return g.qt_global ? quadtree_node_bytes(g.node_cap) : 0; }

int launch_quadtree(dsx_ctx* ctx, int n) {
    StageTimer _t(ctx, 2);
    const ShapePlan& P = ctx->plan;
    QtArgs A;
    size_t smem = 0;
    for (int l = 0; l < P.nlevels; l++) {
        A.lv[l] = P.lv[l];
        smem = std::max(smem, quadtree_smem_bytes(P.lv[l], P.lv[l].qt_depth));
    }
    A.nlevels = P.nlevels;
    A.hist = ctx->ws.hist; A.gbest = ctx->ws.gbest; A.deep = ctx->ws.deep;
    A.node_scratch = (uint8_t*)ctx->ws.node_scratch; A.node_scratch_stride = (long long)P.qt_scratch_stride;
    A.xlut = P.d_xlut; A.ylut = P.d_ylut; A.hist_total = P.hist_total;
    A.cell_count = ctx->ws.cell_count; A.stage = ctx->ws.stage;
    A.cand_xy = ctx->ws.cand_xy; A.cand_resp = ctx->ws.cand_resp; A.cand_node = ctx->ws.cand_node;
    A.cand_count = ctx->ws.cand_count;
    A.key_xy = ctx->ws.key_xy; A.key_resp = ctx->ws.key_resp; A.key_count = ctx->ws.key_count;
    A.cells_total = P.cells_total; A.stage_total = P.stage_total; A.cand_total = P.cand_total;
    A.keys_total = P.keys_total;
    A.err_flag = ctx->ws.err_flag;
    DSX_CUDA(cudaMemsetAsync(ctx->ws.deep, 0, sizeof(int32_t) * DSX_MAX_LEVELS * (size_t)n, ctx->stream));
    DSX_CUDA(cudaFuncSetAttribute(quadtree_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DSX_CUDA(cudaFuncSetAttribute(quadtree_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(P.nlevels, n);
    quadtree_kernel<false><<<grid, kThreads, smem, ctx->stream>>>(A);
    DSX_LAUNCH_CHECK();
    quadtree_kernel<true><<<grid, kThreads, smem, ctx->stream>>>(A);      // exits at once unless `deep` was raised
    DSX_LAUNCH_CHECK();
    return DSX_OK;
}

}  // namespace dsx
