#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int box, int nbox, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* buf = reinterpret_cast<float*>(smem + 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(nbox * box * 4) : "memory");
    if (threadIdx.x < nbox)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_addr(buf + threadIdx.x * box)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0 + (int)threadIdx.x * box), "r"(c1), "r"(smem_addr(bar)) : "memory");
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(0), "r"(2000) : "memory");
    for (int i = threadIdx.x; i < nbox * box; i += blockDim.x) out[i] = buf[i];
}
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult res;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &res) != cudaSuccess || res != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 1; }
    EncodeTiled encode = reinterpret_cast<EncodeTiled>(p);
    const int T = 30000, rows = 10;
    std::vector<float> h(size_t(T) * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = float(i % 100003);
    float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 8192 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int box : {256, 128, 64}) {
        CUtensorMap map{};
        const cuuint64_t dims[2] = {T, rows}; const cuuint64_t strides[1] = {cuuint64_t(T) * 4};
        const cuuint32_t boxes[2] = {cuuint32_t(box), 1}; const cuuint32_t elem[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, boxes, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box %d encode -> %d\n", box, int(r));
        for (int c0 : {0, 1, 1237, 29900}) {
            const int nbox = 5, c1 = 3;
            probe<<<1, 64, 128 + nbox * box * 4>>>(map, c0, c1, box, nbox, out);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> got(nbox * box);
            cudaMemcpy(got.data(), out, got.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int i = 0; i < nbox * box; ++i) { float want = (c0 + i < T) ? h[size_t(c1) * T + c0 + i] : 0.0f; if (got[i] != want) ++bad; }
            printf("  c0 %d: %s, mismatches %d\n", c0, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 2;
        }
    }
    return 0;
}
