// Pointwise (1x1) convolution as a tcgen05 tensor-core GEMM for sm_100a.
//
//   D[M, N] = A[M, K] . W[N, K]^T     A = NHWC bf16 activations (K-major), W = bf16 weights (K-major)
//
// One CTA computes a 128 x BLOCK_N output tile (BLOCK_N is a runtime multiple of 16, <= 256):
//   warp 0      TMA producer : cp.async.bulk.tensor 2D loads of the A and W k-blocks (64 bf16 = one
//                              128-byte swizzle row) into a NUM_STAGES-deep shared-memory ring,
//                              completion signalled on mbarriers (expect_tx)
//   warp 1      MMA issuer   : allocates TMEM, one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                              (UMMA 128 x BLOCK_N x 16, fp32 accumulator in TMEM), tcgen05.commit frees
//                              the smem stage / publishes the accumulator
//   warps 2..5  epilogue     : tcgen05.ld the accumulator (thread = output row), add the folded-BN bias,
//                              activation, residual add, convert, store with the head's strided
//                              addressing (PwEpilogue)
// K and N tails are zero-filled by TMA (out-of-bounds box elements), M tails are masked in the epilogue.
// Reference ops replaced: see dn_pwconv in include/demonet_b200.h.
#include <cuda.h>

#include "common.cuh"
#include "pwconv.cuh"

namespace dn {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;                 // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 16;
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr int TC_A_STAGE_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;      // 16 KiB

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, single-CTA, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major
__device__ __forceinline__ uint32_t make_idesc(int umma_m, int umma_n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}

struct __align__(8) TcBarriers {
    uint64_t full[TC_MAX_STAGES];
    uint64_t empty[TC_MAX_STAGES];
    uint64_t tmem_full;
    uint32_t tmem_base;
    uint32_t pad;
};

__global__ void __launch_bounds__(TC_THREADS)
pwconv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, PwEpilogue ep,
                 int M, int K, int N, int block_n, int n_tiles, int num_stages, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ float s_bias[256];
    // carve: [stages x A tile][stages x W tile][barriers]; tiles must be 1024-B aligned for SWIZZLE_128B
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int w_stage_bytes = block_n * TC_BLOCK_K * 2;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + num_stages * TC_A_STAGE_BYTES;
    TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem_w + num_stages * w_stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x % n_tiles, m_tile = blockIdx.x / n_tiles;
    const int m0 = m_tile * TC_BLOCK_M, n0 = n_tile * block_n;
    const int num_k_blocks = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, (uint32_t)tmem_cols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = bars->tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            const uint32_t stage_bytes = (uint32_t)(TC_A_STAGE_BYTES + w_stage_bytes);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % num_stages;
                const uint32_t round = (uint32_t)(kb / num_stages);
                mbar_wait(&bars->empty[s], (round & 1u) ^ 1u);
                mbar_expect_tx(&bars->full[s], stage_bytes);
                tma_load_2d(smem_a + s * TC_A_STAGE_BYTES, &tmap_a, &bars->full[s], kb * TC_BLOCK_K, m0);
                tma_load_2d(smem_w + s * w_stage_bytes, &tmap_w, &bars->full[s], kb * TC_BLOCK_K, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(TC_BLOCK_M, block_n);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
            const int s = kb % num_stages;
            const uint32_t round = (uint32_t)(kb / num_stages);
            mbar_wait(&bars->full[s], round & 1u);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint64_t da = make_smem_desc(smem_u32(smem_a + s * TC_A_STAGE_BYTES));
                const uint64_t dw = make_smem_desc(smem_u32(smem_w + s * w_stage_bytes));
#pragma unroll
                for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                    // advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in (addr >> 4) units
                    umma_f16(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&bars->empty[s]);                       // frees this smem stage when the MMAs retire
                if (kb == num_k_blocks - 1) umma_commit(&bars->tmem_full);   // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int m = m0 + row;
        const int n_valid = min(block_n, N - n0);
        for (int i = threadIdx.x - 64; i < block_n; i += 128) s_bias[i] = (i < n_valid) ? __ldg(ep.bias + n0 + i) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&bars->tmem_full, 0);
        tcgen05_fence_after();
        const long long out_row = (m < M) ? ep.row_offset(m) : 0;
        const __nv_bfloat16* res_row = ep.residual ? ep.residual + (long long)(m < M ? m : 0) * N : nullptr;
        for (int c0 = 0; c0 < n_valid; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (m < M) {
                const int n = n0 + c0;
                const int cnt = min(16, N - n);
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = apply_act(__uint_as_float(v[i]) + s_bias[c0 + i], ep.act);
                if (res_row) {
                    const __nv_bfloat16* rp = res_row + n;
                    if (cnt == 16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
                        float r[16];
                        unpack8(__ldg(reinterpret_cast<const uint4*>(rp)), r);
                        unpack8(__ldg(reinterpret_cast<const uint4*>(rp) + 1), r + 8);
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] += r[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < cnt) f[i] += __bfloat162float(rp[i]);
                    }
                }
                if (ep.out_fp32) {
                    float* dst = reinterpret_cast<float*>(ep.y) + out_row + n;
                    if (cnt == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            reinterpret_cast<float4*>(dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                    } else if (cnt == 16 && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) reinterpret_cast<float2*>(dst)[i] = make_float2(f[2 * i], f[2 * i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < cnt) dst[i] = f[i];
                    }
                } else {
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.y) + out_row + n;
                    if (cnt == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                        reinterpret_cast<uint4*>(dst)[0] = pack8(f);
                        reinterpret_cast<uint4*>(dst)[1] = pack8(f + 8);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < cnt) dst[i] = __float2bfloat16_rn(f[i]);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
    }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2D bf16 tensor map over a row-major [rows, cols] matrix, box = [box_rows, 64 cols], SWIZZLE_128B
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    DN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, DN_ERR_INVALID, "GEMM operand must be 16-byte aligned");
    DN_REQUIRE(cols % 8 == 0, DN_ERR_UNSUPPORTED, "GEMM K must be a multiple of 8 (got %lld)", cols);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d", (int)r,
               rows, cols, box_rows);
    return DN_OK;
}

void pwconv_tc_plan(int K, int N, int* block_n, int* n_tiles, int* stages, int* tmem_cols, size_t* smem_bytes) {
    const int nt = (N + 255) / 256;
    int bn = (N + nt - 1) / nt;
    bn = (bn + 15) & ~15;
    const int kb = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;
    const int st = kb < TC_MAX_STAGES ? kb : TC_MAX_STAGES;
    int cols = 32;
    while (cols < bn) cols <<= 1;
    *block_n = bn;
    *n_tiles = nt;
    *stages = st;
    *tmem_cols = cols;
    *smem_bytes = 1024 + (size_t)st * (TC_A_STAGE_BYTES + (size_t)bn * TC_BLOCK_K * 2) + sizeof(TcBarriers);
}

int pwconv_tc_launch(const CUtensorMap& ta, const CUtensorMap& tw, const PwEpilogue& ep, int M, int K, int N,
                     cudaStream_t stream) {
    int bn, nt, st, cols;
    size_t smem;
    pwconv_tc_plan(K, N, &bn, &nt, &st, &cols, &smem);
    static size_t configured = 0;
    if (smem > configured) {
        DN_CHECK_CUDA(cudaFuncSetAttribute(pwconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        configured = 200 * 1024;
    }
    const long long tiles = (long long)ceil_div(M, TC_BLOCK_M) * nt;
    DN_REQUIRE(tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "GEMM too large");
    pwconv_tc_kernel<<<(unsigned)tiles, TC_THREADS, smem, stream>>>(ta, tw, ep, M, K, N, bn, nt, st, cols);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

int pwconv_tc(const void* x, const void* w, const PwEpilogue& ep, int M, int K, int N, cudaStream_t stream) {
    int bn, nt, st, cols;
    size_t smem;
    pwconv_tc_plan(K, N, &bn, &nt, &st, &cols, &smem);
    CUtensorMap ta, tw;
    int rc = make_tmap_bf16_2d(&ta, x, M, K, TC_BLOCK_M);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tw, w, N, K, bn);
    if (rc) return rc;
    return pwconv_tc_launch(ta, tw, ep, M, K, N, stream);
}

}  // namespace dn

using namespace dn;

extern "C" int dn_pwconv(const void* x, const void* w, const float* bias, const void* residual, void* y, int M, int K,
                         int N, int act, int out_fp32, int hw, int64_t out_batch_stride, int64_t out_row_stride, int impl,
                         void* stream_) {
    DN_REQUIRE(x && w && bias && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(M > 0 && K > 0 && N > 0 && hw > 0, DN_ERR_INVALID, "bad GEMM shape");
    DN_REQUIRE(K % 8 == 0, DN_ERR_UNSUPPORTED, "GEMM K (input channels) must be a multiple of 8 (got %d)", K);
    DN_REQUIRE(impl == 0 || impl == 1, DN_ERR_INVALID, "impl must be 0 (tcgen05) or 1 (SIMT self-check)");
    PwEpilogue ep;
    ep.bias = bias;
    ep.residual = (const __nv_bfloat16*)residual;
    ep.y = y;
    ep.N = N;
    ep.act = act;
    ep.out_fp32 = out_fp32;
    ep.hw = hw;
    ep.out_batch_stride = out_batch_stride;
    ep.out_row_stride = out_row_stride;
    cudaStream_t s = (cudaStream_t)stream_;
    return impl == 0 ? pwconv_tc(x, w, ep, M, K, N, s) : pwconv_simt(x, w, ep, M, K, N, s);
}
