#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, int x, int y, uint8_t *out)
{
    __shared__ alignas(128) uint8_t smem_buffer[32][64];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}
int main(int argc, char **argv)
{
    const int dt = argc > 1 ? atoi(argv[1]) : 0, l2 = argc > 2 ? atoi(argv[2]) : 0;
    const int P = 640, H = 480;
    std::vector<uint8_t> h((size_t)P * H);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d, *out;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 2048);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m;
    const cuuint64_t dims[2] = { dt ? P / 4u : (unsigned)P, H }; const cuuint64_t strides[1] = { P };
    const cuuint32_t box[2] = { dt ? 16u : 64u, 32 }, es[2] = { 1, 1 };
    CUresult r = ((EncodeTiledFn)p)(&m, dt ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    k<<<1, 64>>>(m, dt ? 25 : 101, 57, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint8_t> o(2048);
    cudaMemcpy(o.data(), out, 2048, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < 32; ++r2) for (int c = 0; c < 64; ++c) bad += o[r2 * 64 + c] != h[(size_t)(57 + r2) * P + (dt ? 100 : 101) + c];
    printf("err=%s mismatches=%d\n", cudaGetErrorString(e), bad);
    return 0;
}
