// Stand-alone check of the TMA strip load used by fast.cu (tensor of 32-bit words, 3-D, box = SP/4 x BH x 1).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x4, int y, int z, int SP, int BH, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar_;
    const uint32_t bar = smem_u32(&bar_);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SP * BH)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(smem)), "l"(&tmap), "r"(x4), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < SP * BH; i += blockDim.x) out[i] = smem[i];
}
__global__ void k_ptr(const CUtensorMap* tmap, int x4, int y, int z, int SP, int BH, uint8_t* out, int rank) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar_;
    const uint32_t bar = smem_u32(&bar_);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SP * BH)) : "memory");
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(smem)), "l"(tmap), "r"(x4), "r"(y), "r"(z), "r"(bar) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(smem)), "l"(tmap), "r"(x4), "r"(y), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < SP * BH; i += blockDim.x) out[i] = smem[i];
}
__global__ void k_bulk(const uint8_t* src, int pitch, int SP, int BH, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar_;
    const uint32_t bar = smem_u32(&bar_);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SP * BH)) : "memory");
    __syncthreads();
    if (threadIdx.x < BH)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + threadIdx.x * SP)), "l"(src + (size_t)threadIdx.x * pitch + 16), "r"((uint32_t)SP), "r"(bar) : "memory");
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < SP * BH; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int rows = 100, pitch = 2000, n = 2, SP = 272, BH = 36;
    std::vector<uint8_t> h((size_t)rows * pitch * n);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, SP * BH);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    const cuuint64_t dims[3] = {(cuuint64_t)(pitch / 4), (cuuint64_t)rows, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
    const cuuint32_t box[3] = {(cuuint32_t)(SP / 4), (cuuint32_t)BH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    // x4 = first 32-bit word of the box (4 * x4 bytes into the row), x8 = first byte of the box for the u8 maps
    const int x4 = argc > 2 ? atoi(argv[2]) : 3, x8 = argc > 3 ? atoi(argv[3]) : 12;
    const int y = 80, z = 1;    // rows 80..115: the last 16 are outside the tensor
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SP * BH + 1024);
    if (mode == 0) k<<<1, 256, SP * BH, 0>>>(tmap, x4, y, z, SP, BH, o);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<uint8_t> g(SP * BH);
    cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BH; r2++) for (int c = 0; c < SP; c++) {
        const int gy = y + r2, gx = 4 * x4 + c;
        const uint8_t want = (gy < rows && gx < pitch) ? h[(size_t)z * pitch * rows + (size_t)gy * pitch + gx] : 0;
        if (g[r2 * SP + c] != want) bad++;
    }
    printf("mismatches: %d\n", bad);
    // variant: the same map read from global memory
    CUtensorMap* dmap; cudaMalloc(&dmap, sizeof(tmap)); cudaMemcpy(dmap, &tmap, sizeof(tmap), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_ptr, cudaFuncAttributeMaxDynamicSharedMemorySize, SP * BH + 1024);
    cudaMemset(o, 0, SP * BH);
    if (mode == 1) k_ptr<<<1, 256, SP * BH, 0>>>(dmap, x4, y, z, SP, BH, o, 3);
    e = cudaDeviceSynchronize();
    printf("kernel (map in global memory, 3d u32): %s\n", cudaGetErrorString(e));
    // variant: 2-D u32 map
    CUtensorMap t2; memset(&t2, 0, sizeof(t2));
    r = ((EncodeTiledFn)p)(&t2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 2d: %d\n", (int)r);
    cudaMemcpy(dmap, &t2, sizeof(t2), cudaMemcpyHostToDevice);
    if (mode == 2) k_ptr<<<1, 256, SP * BH, 0>>>(dmap, x4, y, z, SP, BH, o, 2);
    e = cudaDeviceSynchronize();
    printf("kernel (2d u32): %s\n", cudaGetErrorString(e));
    // variant: 2-D u8 map, box 256 x BH
    CUtensorMap t3; memset(&t3, 0, sizeof(t3));
    const cuuint64_t dims8[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
    const cuuint32_t box8[2] = {256, (cuuint32_t)BH};
    r = ((EncodeTiledFn)p)(&t3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims8, strides, box8, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 2d u8: %d\n", (int)r);
    cudaMemcpy(dmap, &t3, sizeof(t3), cudaMemcpyHostToDevice);
    if (mode == 3) k_ptr<<<1, 256, 256 * BH, 0>>>(dmap, x8, y, z, 256, BH, o, 2);
    if (mode == 4) k_ptr<<<1, 256, 256 * BH, 0>>>(dmap, x8, 10, z, 256, BH, o, 2);
    e = cudaDeviceSynchronize();
    printf("kernel (2d u8): %s\n", cudaGetErrorString(e));
    if (mode == 5) {
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, SP * BH + 1024);
        k_bulk<<<1, 256, SP * BH, 0>>>(d, pitch, SP, BH, o);
        e = cudaDeviceSynchronize();
        printf("kernel (bulk rows): %s\n", cudaGetErrorString(e));
        cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
        bad = 0;
        for (int r2 = 0; r2 < BH; r2++) for (int c = 0; c < SP; c++) if (g[r2 * SP + c] != h[(size_t)r2 * pitch + 16 + c]) bad++;
        printf("bulk mismatches: %d\n", bad);
    }
    return 0;
}
