// SIMT (CUDA-core) pointwise GEMM: bring-up and on-device self-check twin of the tcgen05 kernel in
// pwconv_tc.cu.  Same arguments, same epilogue, same bf16 inputs / fp32 accumulation; it exists so
// that the tensor-core kernel can be compared against an independent implementation ON THE GPU at
// full problem sizes (tests/test_pwconv_gpu.py).  It is NOT selected on the product path
// (dn_model_desc.gemm_impl defaults to 0 = tcgen05).
#include "common.cuh"
#include "pwconv.cuh"

namespace dn {

constexpr int ST_BM = 64, ST_BN = 64, ST_BK = 16;

__global__ void __launch_bounds__(256)
pwconv_simt_kernel(const dn_half_t* __restrict__ x, const dn_half_t* __restrict__ w, PwEpilogue ep, int M,
                   int K, int N) {
    __shared__ float As[ST_BK][ST_BM + 1];
    __shared__ float Ws[ST_BK][ST_BN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.x * ST_BM, n0 = blockIdx.y * ST_BN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += ST_BK) {
        for (int i = threadIdx.x; i < ST_BM * ST_BK; i += 256) {
            const int r = i / ST_BK, c = i % ST_BK;
            const int m = m0 + r, k = k0 + c;
            As[c][r] = (m < M && k < K) ? half_to_float(x[(long long)m * K + k]) : 0.f;
            const int n = n0 + r;
            Ws[c][r] = (n < N && k < K) ? half_to_float(w[(long long)n * K + k]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < ST_BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) ep.store(m, n, acc[i][j]);
        }
    }
}

int pwconv_simt(const void* x, const void* w, const PwEpilogue& ep, int M, int K, int N, cudaStream_t stream) {
    dim3 grid(ceil_div(M, ST_BM), ceil_div(N, ST_BN));
    pwconv_simt_kernel<<<grid, 256, 0, stream>>>((const dn_half_t*)x, (const dn_half_t*)w, ep, M, K, N);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

}  // namespace dn
