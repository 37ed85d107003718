// Minimal TMA 2D load probe: which descriptor / addressing variants work on this box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int BH>
__global__ void k_probe(const __grid_constant__ CUtensorMap map, const CUtensorMap* gmap, int use_global, int x, int y,
                        uint32_t* out) {
    __shared__ __align__(128) uint32_t buf[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH * 4) : "memory");
        const CUtensorMap* m = use_global ? gmap : &map;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf)), "l"(m), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BW, int BH>
void run(EncodeTiledFn fn, uint32_t* d_in, int W, int H, int use_global, CUtensorMapL2promotion l2) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    cuuint32_t box[2] = {BW, BH};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap* gmap;
    cudaMalloc(&gmap, sizeof(map));
    cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    uint32_t* d_out;
    cudaMalloc(&d_out, BW * BH * 4);
    cudaMemset(d_out, 0xff, BW * BH * 4);
    k_probe<BW, BH><<<1, 128>>>(map, gmap, use_global, -5, -2, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint32_t> h(BW * BH);
    cudaMemcpy(h.data(), d_out, BW * BH * 4, cudaMemcpyDeviceToHost);
    printf("box %dx%d W=%d global_desc=%d l2=%d encode=%d -> %s ; out[2*BW+5]=%u (want %u) out[0]=%u\n", BW, BH, W, use_global,
           (int)l2, (int)r, cudaGetErrorString(e), h[2 * BW + 5], 0u * W + 0u + 1000u, h[0]);
    cudaFree(gmap);
    cudaFree(d_out);
}

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int W = 96, H = 96;
    std::vector<uint32_t> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = 1000 + i;
    uint32_t* d_in;
    cudaMalloc(&d_in, W * H * 4);
    cudaMemcpy(d_in, h.data(), W * H * 4, cudaMemcpyHostToDevice);
    run<64, 32>(fn, d_in, W, H, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    run<64, 32>(fn, d_in, W, H, 1, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    run<72, 36>(fn, d_in, W, H, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    run<72, 36>(fn, d_in, W, H, 1, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    run<72, 36>(fn, d_in, W, H, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    run<72, 36>(fn, d_in, W, H, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    return 0;
}
