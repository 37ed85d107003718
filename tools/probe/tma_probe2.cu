#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define WAIT(bar) asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)) : "memory")

__global__ void k_bulk1d(const uint32_t* src, uint32_t* out) {
    __shared__ __align__(128) uint32_t buf[1024];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(buf)), "l"(src), "r"(4096), "r"(smem_u32(&bar)) : "memory");
    }
    WAIT(&bar);
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}

__global__ void k_tensor(const __grid_constant__ CUtensorMap map, int x, int y, uint32_t* out) {
    __shared__ __align__(128) uint32_t buf[64 * 32];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(64 * 32 * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf)), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    WAIT(&bar);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) out[i] = buf[i];
}

__global__ void k_tensor_g(const CUtensorMap* map, int x, int y, uint32_t* out) {
    __shared__ __align__(128) uint32_t buf[64 * 32];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(64 * 32 * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    WAIT(&bar);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 256, H = 128;
    std::vector<uint32_t> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = 1000 + i;
    uint32_t *d_in, *d_out;
    cudaMalloc(&d_in, W * H * 4);
    cudaMalloc(&d_out, 64 * 32 * 4);
    cudaMemcpy(d_in, h.data(), W * H * 4, cudaMemcpyHostToDevice);
    std::vector<uint32_t> o(64 * 32);
    if (which == 0) {
        k_bulk1d<<<1, 128>>>(d_in, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(o.data(), d_out, 4096, cudaMemcpyDeviceToHost);
        printf("bulk1d: %s out[5]=%u (want 1005)\n", cudaGetErrorString(e), o[5]);
        return 0;
    }
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap map;
    cuuint64_t dims[2] = {W, H};
    cuuint64_t strides[1] = {W * 4};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d query=%d\n", (int)r, (int)q);
    int x = argc > 2 ? atoi(argv[2]) : 64, y = argc > 3 ? atoi(argv[3]) : 32;
    if (which == 1 || which == 3) {
        k_tensor<<<1, 128>>>(map, x, y, d_out);
    } else {
        CUtensorMap* gmap;
        cudaMalloc(&gmap, sizeof(map));
        cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
        k_tensor_g<<<1, 128>>>(gmap, x, y, d_out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(o.data(), d_out, 64 * 32 * 4, cudaMemcpyDeviceToHost);
    printf("tensor variant %d (x=%d,y=%d): %s out[0]=%u out[2*64+5]=%u\n", which, x, y, cudaGetErrorString(e), o[0], o[2 * 64 + 5]);
    return 0;
}
