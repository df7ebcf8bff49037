// Probe: which fp32 3-D TMA boxes does the hardware accept?  usage: tma_probe W H P box0 box1 box2 x y z
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int n, int x, int y, int z) {
    extern __shared__ __align__(128) float t[];
    __shared__ uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(t);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)(n * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(d),
                     "l"(&tm), "r"(b), "r"(x), "r"(y), "r"(z)
                     : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(b), "r"(0u) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = t[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int W = atoi(argv[1]), H = atoi(argv[2]), P = atoi(argv[3]), b0 = atoi(argv[4]), b1 = atoi(argv[5]), b2 = atoi(argv[6]);
    int x = atoi(argv[7]), y = atoi(argv[8]), z = atoi(argv[9]);
    float* g; cudaMalloc(&g, (size_t)W * H * P * 4);
    float* h = (float*)malloc((size_t)W * H * P * 4);
    for (int i = 0; i < W * H * P; ++i) h[i] = (float)i;
    cudaMemcpy(g, h, (size_t)W * H * P * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P}; cuuint64_t gs[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t bx[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((PFN)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int n = b0 * b1 * b2; float* out; cudaMalloc(&out, n * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<1, 128, n * 4 + 128>>>(tm, out, n, x, y, z);
    cudaError_t e = cudaDeviceSynchronize();
    float* ho = (float*)malloc(n * 4); cudaMemcpy(ho, out, n * 4, cudaMemcpyDeviceToHost);
    // expected value of element (i0,i1,i2) = ((z+i2)*H + (y+i1))*W + (x+i0) when inside, else 0
    int bad = 0;
    for (int i2 = 0; i2 < b2; ++i2) for (int i1 = 0; i1 < b1; ++i1) for (int i0 = 0; i0 < b0; ++i0) {
        int gx = x + i0, gy = y + i1, gz = z + i2;
        float exp = (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < P) ? (float)((gz * H + gy) * W + gx) : 0.f;
        if (ho[(i2 * b1 + i1) * b0 + i0] != exp) ++bad;
    }
    printf("W=%d H=%d P=%d box=%d,%d,%d at %d,%d,%d: encode=%d run=%s mismatches=%d\n", W, H, P, b0, b1, b2, x, y, z, (int)r,
           cudaGetErrorString(e), bad);
    return 0;
}
