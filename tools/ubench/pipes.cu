// Micro-benchmark: issue rates of the integer min/max flavours on sm_100a and whether fp16x2 min/max overlaps with them.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed) {
    unsigned a[8], b[8];
    __half2 h[8], g[8];
    for (int i = 0; i < 8; i++) {
        a[i] = threadIdx.x * 2654435761u + i * 40503u + seed; b[i] = a[i] ^ 0x12345678u;
        h[i] = __halves2half2(__ushort_as_half((unsigned short)(a[i] & 0x3fff)), __ushort_as_half((unsigned short)((a[i] >> 16) & 0x3fff)));
        g[i] = __halves2half2(__ushort_as_half((unsigned short)(b[i] & 0x3fff)), __ushort_as_half((unsigned short)((b[i] >> 16) & 0x3fff)));
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]);
            if (MODE == 1) a[i] = __vminu2(a[i], b[i] + it);
            if (MODE == 2) h[i] = __hmin2(h[i], __hadd2(g[i], g[(i + 1) & 7]));   // hadd keeps it from collapsing
            if (MODE == 3) h[i] = __hmin2(h[i], g[(i + it) & 7]);
            if (MODE == 4) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); h[i] = __hmin2(h[i], g[(i + it) & 7]); }
            if (MODE == 5) a[i] = __vimin3_s32(a[i], b[i], a[(i + 1) & 7]);
            if (MODE == 6) a[i] = __funnelshift_r(a[i], b[i], 8) ^ a[(i + 1) & 7];
            if (MODE == 7) a[i] = a[i] * 3u + b[i];                                    // IMAD (fma pipe)
            if (MODE == 8) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = b[i] * 3u + a[(i + 2) & 7]; }
            if (MODE == 9) a[i] = __vabsdiffu4(a[i], b[i]) + a[(i + 1) & 7];
            if (MODE == 10) a[i] = __byte_perm(a[i], b[i], 0x6543) ;
            if (MODE == 11) a[i] = a[i] & b[i] ^ a[(i + 1) & 7];
            if (MODE == 12) { float x = __uint_as_float(a[i]), y = __uint_as_float(b[i]), z = __uint_as_float(a[(i + 1) & 7]), r;
                              asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x), "f"(y), "f"(z)); a[i] = __float_as_uint(r); }
            if (MODE == 13) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]);
                              float x = __uint_as_float(b[i]), y = __uint_as_float(b[(i + 3) & 7]), z = __uint_as_float(b[(i + 1) & 7]), r;
                              asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x), "f"(y), "f"(z)); b[i] = __float_as_uint(r); }
            if (MODE == 14) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = (b[i] << 8) + a[(i + 2) & 7]; }   // shift as IMAD?
            if (MODE == 15) a[i] = __popc(a[i] ^ b[i]) + a[(i + 1) & 7];
            if (MODE == 16) { a[i] = __popc(a[i] ^ b[i]) + a[(i + 1) & 7]; b[i] = b[i] * 3u + a[(i + 2) & 7]; b[i] = (b[i] & a[i]) ^ a[(i+3)&7]; }
            if (MODE == 17) { double x = __hiloint2double(a[i], b[i]); x = __dadd_rn(__dmul_rn(x, x), 1.0); a[i] = __double2hiint(x); b[i] = __double2loint(x); }
        }
    }
    unsigned r = 0;
    for (int i = 0; i < 8; i++) r += a[i] + b[i] + __half_as_ushort(__low2half(h[i])) + __half_as_ushort(__high2half(h[i]));
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
    // per SMSP: 148*4 SMSPs; clock ~1.9 GHz
    printf("%-34s %7.3f ms  %.3f warp-instr/clk/SMSP (at 1.9 GHz)\n", name, ms, warp_instr / (ms * 1e-3) / (148.0 * 4) / 1.9e9);
}
int main() {
    run<0>("VIMNMX3.U16x2", 8); run<1>("VIMNMX.U16x2 (+IADD)", 16); run<2>("HMNMX2 + HADD2", 16); run<3>("HMNMX2", 8);
    run<4>("VIMNMX3.U16x2 + HMNMX2", 16); run<5>("VIMNMX3.S32", 8); run<6>("SHF + LOP3", 16); run<7>("IMAD", 8);
    run<8>("VIMNMX3.U16x2 + IMAD", 16); run<9>("VABSDIFF4 + IADD", 16);
    run<10>("PRMT", 8); run<11>("LOP3", 8); run<12>("FMNMX3 (min.f32 3-in)", 8); run<13>("VIMNMX3.U16x2 + FMNMX3", 16);
    run<14>("VIMNMX3.U16x2 + (x<<8)+y", 16); run<15>("POPC + LOP3 + IADD", 24); run<16>("POPC+LOP3+IADD + IMAD+LOP3", 40);
    run<17>("DMUL + DADD", 16);
    return 0;
}
