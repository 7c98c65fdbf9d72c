// tma_probe3.cu -- does the TMA engine work on this pool?  One test per process (a faulting kernel kills the context):
//   ./tma_probe3 bulk      1-D cp.async.bulk.shared::cluster.global + mbarrier complete_tx (raw PTX)
//   ./tma_probe3 tensor    2-D cp.async.bulk.tensor.2d (raw PTX, cuTensorMapEncodeTiled through cudaGetDriverEntryPoint)
//   ./tma_probe3 prefetch  cp.async.bulk.prefetch.tensor (no shared-memory destination at all)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe3 tools/tma_probe3.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

__global__ void k_bulk(const uint8_t *src, int pitch, uint8_t *out)
{
    __shared__ alignas(128) uint8_t buf[32][64];
    __shared__ alignas(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 32 * 64);
        for (int r = 0; r < 32; ++r)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(&buf[r][0])), "l"(src + (size_t)(57 + r) * pitch + 96), "r"(64), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = (&buf[0][0])[i];
}

__global__ void k_tensor(const __grid_constant__ CUtensorMap map, int x, int y, uint8_t *out)
{
    __shared__ alignas(128) uint8_t buf[32][64];
    __shared__ alignas(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 32 * 64);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(&buf[0][0])), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = (&buf[0][0])[i];
}

__global__ void k_prefetch(const __grid_constant__ CUtensorMap map, int x, int y, uint8_t *out)
{
    if (threadIdx.x == 0)
        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&map), "r"(x), "r"(y) : "memory");
    out[threadIdx.x] = (uint8_t)threadIdx.x;
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "bulk";
    const int P = 640, H = 480;
    std::vector<uint8_t> h((size_t)P * H);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d, *out;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 2048); cudaMemset(out, 0, 2048);
    int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("mode=%s device=%s cc=%d.%d driver_api=%d runtime=%d\n", mode, prop.name, prop.major, prop.minor, drv, rt);
    int xoff = 96;
    if (!strcmp(mode, "bulk")) {
        k_bulk<<<1, 64>>>(d, P, out);
    } else {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        printf("cudaGetDriverEntryPoint: %s, query=%d, fn=%p\n", cudaGetErrorString(ge), (int)q, p);
        alignas(64) CUtensorMap m;
        const cuuint64_t dims[2] = { (cuuint64_t)P, (cuuint64_t)H }; const cuuint64_t strides[1] = { (cuuint64_t)P };
        const cuuint32_t box[2] = { 64, 32 }, es[2] = { 1, 1 };
        CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("cuTensorMapEncodeTiled -> %d\n", (int)r);
        xoff = 101;
        if (!strcmp(mode, "tensor")) k_tensor<<<1, 64>>>(m, xoff, 57, out);
        else k_prefetch<<<1, 64>>>(m, xoff, 57, out);
    }
    cudaError_t le = cudaGetLastError();
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint8_t> o(2048);
    cudaMemcpy(o.data(), out, 2048, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (strcmp(mode, "prefetch"))
        for (int r2 = 0; r2 < 32; ++r2) for (int c = 0; c < 64; ++c) bad += o[r2 * 64 + c] != h[(size_t)(57 + r2) * P + xoff + c];
    printf("launch=%s sync=%s mismatches=%d\n", cudaGetErrorString(le), cudaGetErrorString(e), bad);
    return 0;
}
