// minimal TMA 3-D bulk copy probe (development aid): which box shapes / coordinates does cp.async.bulk.tensor.3d accept here?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

template <int BX, int BY, int BZ>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x0, int y0, int z0, uint32_t *out)
{
    extern __shared__ __align__(128) uint32_t box[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        if (leader) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(box);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"((uint32_t)(BX * BY * BZ * 4)) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x0), "r"(y0), "r"(z0), "r"(bar_addr) : "memory");
        }
    }
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar_addr), "r"(0u) : "memory");
    if (threadIdx.x == 0) { out[0] = box[0]; out[1] = box[(1 * BY + 2) * BX + 3]; }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BX, int BY, int BZ>
void run(EncodeFn enc, uint32_t *vol, int D, int x0, int y0, int z0, uint32_t *out)
{
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)D, (cuuint64_t)D}, strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)D * D * 4};
    const cuuint32_t box[3] = {BX, BY, BZ}, es[3] = {1, 1, 1};
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, vol, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(probe<BX, BY, BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, BX * BY * BZ * 4);
    cudaMemset(out, 0xff, 8);
    probe<BX, BY, BZ><<<1, 256, BX * BY * BZ * 4>>>(m, x0, y0, z0, out);
    const cudaError_t e = cudaDeviceSynchronize();
    uint32_t h[2] = {0, 0};
    if (e == cudaSuccess) cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    const long long want0 = (x0 >= 0 && y0 >= 0 && z0 >= 0) ? (long long)x0 + (long long)D * (y0 + (long long)D * z0) : -1;
    printf("box %dx%dx%d at (%d,%d,%d): encode %d, run %s, box[0]=%u (want %lld) box[1,2,3]=%u (want %lld)\n", BX, BY, BZ, x0, y0, z0, (int)r, cudaGetErrorString(e), h[0], want0,
           h[1], (long long)(x0 + 3) + (long long)D * ((y0 + 2) + (long long)D * (z0 + 1)));
    if (e != cudaSuccess) { cudaDeviceReset(); exit(1); }
}

int main(int argc, char **argv)
{
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    const int D = 128;
    std::vector<uint32_t> h((size_t)D * D * D);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i;
    uint32_t *vol, *out;
    cudaMalloc(&vol, h.size() * 4); cudaMalloc(&out, 64);
    cudaMemcpy(vol, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    if (which == 0) run<32, 8, 8>(enc, vol, D, 0, 0, 0, out);
    if (which == 1) run<32, 16, 32>(enc, vol, D, 4, 8, 16, out);
    if (which == 2) run<40, 16, 32>(enc, vol, D, 4, 8, 16, out);
    if (which == 3) run<64, 16, 16>(enc, vol, D, 5, 8, 16, out);
    if (which == 4) run<32, 16, 32>(enc, vol, D, -3, -2, 100, out);
    if (which == 5) run<48, 16, 24>(enc, vol, D, 4, 8, 16, out);
    return 0;
}
