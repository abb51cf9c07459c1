#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed) {
    unsigned a0 = threadIdx.x * 2654435761u + seed, a1 = a0 ^ 0x1234567u, a2 = a0 * 3, a3 = a0 * 5, a4 = a0 * 7, a5 = a0 * 11, a6 = a0 * 13, a7 = a0 * 17;
    __half2 h0 = __floats2half2_rn((float)(a0 & 255), (float)(a1 & 255)), h1 = __floats2half2_rn((float)(a2 & 255), (float)(a3 & 255)),
            h2 = __floats2half2_rn((float)(a4 & 255), (float)(a5 & 255)), h3 = __floats2half2_rn((float)(a6 & 255), (float)(a7 & 255)),
            h4 = __hadd2(h0, h1), h5 = __hadd2(h1, h2), h6 = __hadd2(h2, h3), h7 = __hadd2(h3, h0);
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 2) {   // 8 independent 2-input half2 min/max
            h0 = __hmin2(h0, h1); h1 = __hmax2(h1, h2); h2 = __hmin2(h2, h3); h3 = __hmax2(h3, h4);
            h4 = __hmin2(h4, h5); h5 = __hmax2(h5, h6); h6 = __hmin2(h6, h7); h7 = __hmax2(h7, h0);
        }
        if (MODE == 1 || MODE == 2) {   // 8 VIMNMX3.U16x2
            a0 = __vimin3_u16x2(a0, a1, a2); a1 = __vimax3_u16x2(a1, a2, a3); a2 = __vimin3_u16x2(a2, a3, a4); a3 = __vimax3_u16x2(a3, a4, a5);
            a4 = __vimin3_u16x2(a4, a5, a6); a5 = __vimax3_u16x2(a5, a6, a7); a6 = __vimin3_u16x2(a6, a7, a0); a7 = __vimax3_u16x2(a7, a0, a1);
        }
        if (MODE == 3 || MODE == 4) {   // 8 three-input half2 min (nested -> compiler may fuse to VHMNMX)
            h0 = __hmin2(h0, __hmin2(h1, h2)); h1 = __hmax2(h1, __hmax2(h2, h3)); h2 = __hmin2(h2, __hmin2(h3, h4)); h3 = __hmax2(h3, __hmax2(h4, h5));
            h4 = __hmin2(h4, __hmin2(h5, h6)); h5 = __hmax2(h5, __hmax2(h6, h7)); h6 = __hmin2(h6, __hmin2(h7, h0)); h7 = __hmax2(h7, __hmax2(h0, h1));
        }
        if (MODE == 4) {
            a0 = __vimin3_u16x2(a0, a1, a2); a1 = __vimax3_u16x2(a1, a2, a3); a2 = __vimin3_u16x2(a2, a3, a4); a3 = __vimax3_u16x2(a3, a4, a5);
            a4 = __vimin3_u16x2(a4, a5, a6); a5 = __vimax3_u16x2(a5, a6, a7); a6 = __vimin3_u16x2(a6, a7, a0); a7 = __vimax3_u16x2(a7, a0, a1);
        }
        if (MODE == 5) {   // HFMA2.relu based max: y + relu(x - y)
            const __half2 one = __floats2half2_rn(1.f, 1.f);
            h0 = __hadd2(h1, __hfma2_relu(h0, one, __hneg2(h1))); h2 = __hadd2(h3, __hfma2_relu(h2, one, __hneg2(h3)));
            h4 = __hadd2(h5, __hfma2_relu(h4, one, __hneg2(h5))); h6 = __hadd2(h7, __hfma2_relu(h6, one, __hneg2(h7)));
            h1 = __hadd2(h2, __hfma2_relu(h1, one, __hneg2(h2))); h3 = __hadd2(h4, __hfma2_relu(h3, one, __hneg2(h4)));
            h5 = __hadd2(h6, __hfma2_relu(h5, one, __hneg2(h6))); h7 = __hadd2(h0, __hfma2_relu(h7, one, __hneg2(h0)));
        }
    }
    unsigned r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    __half2 hs = __hadd2(__hadd2(__hadd2(h0, h1), __hadd2(h2, h3)), __hadd2(__hadd2(h4, h5), __hadd2(h6, h7)));
    r += __half_as_ushort(__low2half(hs)) + __half_as_ushort(__high2half(hs));
    if (r == 0x12345) out[0] = r;
}
template <int MODE> void run(const char* name, int ops_per_iter) {
    unsigned* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, 2);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * 8 * ITERS * ops_per_iter;
    printf("%-40s %7.3f ms  %.3f ops/clk/SMSP (at 1.965 GHz)\n", name, ms, warp_instr / (ms * 1e-3) / (148.0 * 4) / 1.965e9);
}
int main() {
    run<0>("HMNMX2 (2-in) x8", 8); run<1>("VIMNMX3.U16x2 x8", 8); run<2>("HMNMX2 x8 + VIMNMX3 x8", 16);
    run<3>("half2 3-in min/max x8", 8); run<4>("half2 3-in x8 + VIMNMX3 x8", 16); run<5>("max via HADD2+HFMA2.relu x8 (16 instr)", 8);
    return 0;
}
