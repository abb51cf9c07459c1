// Micro-benchmark: can fp16x2 min/max (HMNMX2) run beside the integer VIMNMX3.U16x2 on sm_100a?
// u8 values encoded as fp16 0x6400|b (= 1024+b) order identically as u16 and as fp16, so a min/max network can be
// split between the two instruction families if they issue to different pipes.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ unsigned hmin2u(unsigned a, unsigned b) { unsigned d; asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) { unsigned d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hfma_relu(unsigned a, unsigned b, unsigned c) { unsigned d; asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed) {
    unsigned a[8], b[8];
    for (int i = 0; i < 8; i++) {
        unsigned v = threadIdx.x * 2654435761u + i * 40503u + seed;
        a[i] = 0x64006400u | (v & 0x00ff00ffu); b[i] = 0x64006400u | ((v >> 8) & 0x00ff00ffu);
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]);
            if (MODE == 1) a[i] = hmin2u(a[i], hmax2u(b[i], a[(i + 1) & 7]));                       // 2 HMNMX2
            if (MODE == 2) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = hmax2u(b[i], a[(i + 3) & 7]); }   // 1 + 1
            if (MODE == 3) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = hmax2u(hmin2u(b[i], a[(i + 3) & 7]), a[(i + 5) & 7]); }   // 1 + 2
            if (MODE == 4) a[i] = hfma_relu(a[i], 0x3c003c00u, b[(i + 1) & 7]);                       // HFMA2.RELU
            if (MODE == 5) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = hfma_relu(b[i], 0x3c003c00u, a[(i + 3) & 7]); }
            if (MODE == 6) a[i] = __vminu2(a[i], __vmaxu2(b[i], a[(i + 1) & 7]));                     // 2 VIMNMX.U16x2
            if (MODE == 7) a[i] = __byte_perm(a[i], b[(i + 1) & 7], 0x6240) + 1;                      // PRMT + IADD
            if (MODE == 8) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = __byte_perm(b[i], a[(i + 3) & 7], 0x6240); }
            if (MODE == 9) { a[i] = hmin2u(a[i], b[(i + 1) & 7]); b[i] = __byte_perm(b[i], a[(i + 3) & 7], 0x6240); }
            if (MODE == 10) { a[i] = hmin2u(a[i], b[(i + 1) & 7]); b[i] = __funnelshift_r(b[i], a[(i + 3) & 7], 8); }
            if (MODE == 11) { a[i] = hmin2u(a[i], b[(i + 1) & 7]); b[i] = b[i] * 5u + a[(i + 3) & 7]; }   // HMNMX2 + IMAD
            if (MODE == 12) a[i] = __umulhi(a[i], 0x01000000u) + b[(i + 1) & 7];                         // IMAD.HI
            if (MODE == 13) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = __umulhi(b[i], 0x01000000u) + a[(i + 3) & 7]; }
            if (MODE == 14) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = __umulhi(b[i], 0x01000000u) + a[(i + 3) & 7] * 256u; }   // 1 ALU + IMAD.HI + IMAD
            if (MODE == 15) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = b[i] * 5u + a[(i + 3) & 7]; }   // VIMNMX3 + IMAD
            if (MODE == 16) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = (b[i] * 5u + a[(i + 3) & 7]) * 3u + a[(i + 5) & 7]; }   // VIMNMX3 + 2 IMAD
        }
    }
    unsigned r = 0;
    for (int i = 0; i < 8; i++) r += a[i] + b[i];
    if (r == 0x12345) out[0] = r;
}
static double g_clk = 1.9e9;
template <int MODE> void run(const char* name, int ops_per_iter) {
    unsigned* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, 2);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * 8 * ITERS * 8 * ops_per_iter;
    printf("%-40s %7.3f ms  %.3f warp-instr/clk/SMSP total (%d instr per step)\n", name, ms, warp_instr / (ms * 1e-3) / (148.0 * 4) / g_clk, ops_per_iter);
}
int main() {
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); if (khz > 0) g_clk = khz * 1e3;
    printf("clock %.0f MHz assumed\n", g_clk / 1e6);
    run<0>("VIMNMX3.U16x2", 1); run<1>("HMNMX2 x2", 2); run<2>("VIMNMX3 + HMNMX2", 2); run<3>("VIMNMX3 + 2 HMNMX2", 3);
    run<4>("HFMA2.RELU", 1); run<5>("VIMNMX3 + HFMA2.RELU", 2); run<6>("VIMNMX.U16x2 x2", 2); run<7>("PRMT + IADD", 2);
    run<8>("VIMNMX3 + PRMT", 2); run<9>("HMNMX2 + PRMT", 2); run<10>("HMNMX2 + SHF", 2); run<11>("HMNMX2 + IMAD", 2);
    run<12>("IMAD.HI", 1); run<13>("VIMNMX3 + IMAD.HI", 2); run<14>("VIMNMX3 + IMAD.HI + IMAD", 3); run<15>("VIMNMX3 + IMAD", 2);
    run<16>("VIMNMX3 + 2 IMAD", 3);
    return 0;
}
