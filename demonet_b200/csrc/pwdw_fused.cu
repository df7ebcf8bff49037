// Fused pointwise-expand (1x1 conv + BN + act) -> depthwise 3x3 stride-2 (+ BN + act) for sm_100a: the expanded
// tensor never leaves the SM.  First instance of DESIGN.md section 8 item 1, for the block where it pays most:
// MobileNetV3 block 2 (16 -> 64 channels at 160x160, then 3x3 s2): the expanded tensor is the largest of the network
// (839 MB per batch of 256, written by the GEMM and read back by the depthwise kernel).
// Reference: InvertedResidual.block[0:2], demonet/models/mobilenetv3.py:75-83 (see dn_pwconv / dn_dwconv).
//
// A persistent CTA walks 8 x 8 output tiles.  Per tile:
//   TMA      one cp.async.bulk.tensor.4d fetches the 17 x 17 x 16-channel input window (NHWC bf16, SWIZZLE_32B: one
//            32-byte row per pixel = the K-major A operand, out-of-image pixels zero-filled), double-buffered
//   MMA      3 x tcgen05.mma (128 pixels x 64 channels x K = 16, kind::f16) -> fp32 accumulators in TMEM (192 columns)
//   drain    all 8 warps: tcgen05.ld -> + bias -> activation -> bf16 -> shared-memory tile [289 pixels][64 channels]
//            (16-byte chunks XOR-swizzled by the pixel index); pixels outside the image are forced to ZERO: they are
//            the depthwise convolution's padding, not act(bias)
//   stencil  thread = 8 channels x 2 output pixels, 9 taps from the shared-memory tile, packed FFMA2, fp32 weights,
//            bias + activation, one 16-byte store per pixel; the MMAs of the next tile run underneath
#include <cuda.h>

#include "common.cuh"
#include "dwconv.cuh"
#include "pwconv.cuh"

namespace dn {

constexpr int FU_TO = 8;                        // output tile edge
constexpr int FU_IT = 2 * FU_TO + 1;            // input tile edge (k = 3, stride 2)
constexpr int FU_PIX = FU_IT * FU_IT;           // 289 pixels
constexpr int FU_MT = (FU_PIX + 127) / 128;     // 3 MMA tiles of 128 pixels
constexpr int FU_K = 16, FU_N = 64;
constexpr int FU_THREADS = 256;
constexpr int FU_A_BYTES = FU_MT * 128 * FU_K * 2;          // 12 KiB per input buffer (289 rows used)
constexpr int FU_EXP_BYTES = FU_PIX * FU_N * 2;             // 36 992 B expanded tile
constexpr int FU_TMEM_COLS = 256;                           // 3 x 64 accumulator columns, power of two

__device__ __forceinline__ uint32_t fu_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fu_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(fu_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
template <int ACT>
__device__ __forceinline__ float fu_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));
    return v;
}
// K-major SWIZZLE_32B operand descriptor: 32-byte rows, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t fu_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}

struct __align__(8) FuBars {
    uint64_t full[2];          // TMA -> MMA: input window landed
    uint64_t mma_done;         // MMA -> drain: accumulators complete
    uint64_t w_full;           // expand weights landed
    uint32_t tmem_base;
    uint32_t pad;
};

template <int ACT_PW, int ACT_DW>
__global__ void __launch_bounds__(FU_THREADS, 2)
pwdw_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const float* __restrict__ b_pw, const float* __restrict__ w_dw, const float* __restrict__ b_dw,
                  uint4* __restrict__ y, int H, int W, int Ho, int Wo, int tiles_x, int tiles_y, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char fu_smem_raw[];
    // aligned by pointer arithmetic on the __shared__ array (not through an integer cast) so that the compiler keeps the
    // shared address space and emits LDS / STS instead of generic loads and stores
    unsigned char* fu_smem = fu_smem_raw + ((1024u - (fu_u32(fu_smem_raw) & 1023u)) & 1023u);
    unsigned char* a_in = fu_smem;                                  // [2][FU_A_BYTES]
    unsigned char* w_s = a_in + 2 * FU_A_BYTES;                     // [64 x 16] bf16, SWIZZLE_32B
    unsigned char* exp_t = w_s + 2048;                              // [289][64] bf16, chunk-swizzled
    float* wdw = reinterpret_cast<float*>(exp_t + ((FU_EXP_BYTES + 127) & ~127));      // [9][64]
    float* bdw = wdw + 9 * FU_N;                                    // [64]
    float* bpw = bdw + FU_N;                                        // [64]
    FuBars* bars = reinterpret_cast<FuBars*>(bpw + FU_N);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fu_u32(&bars->full[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fu_u32(&bars->full[1])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fu_u32(&bars->mma_done)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fu_u32(&bars->w_full)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fu_u32(&bars->tmem_base)),
                     "r"((uint32_t)FU_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 9 * FU_N; i += FU_THREADS) wdw[i] = __ldg(w_dw + i);
    for (int i = threadIdx.x; i < FU_N; i += FU_THREADS) bdw[i] = __ldg(b_dw + i), bpw[i] = __ldg(b_pw + i);
    pdl_trigger();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    auto issue_tile = [&](int t, int buf) {       // one thread: fetch the input window of tile t
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const uint32_t bar = fu_u32(&bars->full[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(FU_PIX * FU_K * 2)) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                "r"(fu_u32(a_in + buf * FU_A_BYTES)),
            "l"(&tmap_x), "r"(bar), "r"(0), "r"(tx * FU_TO * 2 - 1), "r"(ty * FU_TO * 2 - 1), "r"(b)
            : "memory");
    };
    auto issue_mma = [&](int buf) {               // one thread: 3 MMAs over the 384 (289 used) pixel rows
        // instruction descriptor: D = f32, A = B = bf16, K-major both, N = 64, M = 128
        constexpr uint32_t idesc = (1u << 4) | (DN_UMMA_AB_FORMAT << 7) | (DN_UMMA_AB_FORMAT << 10) | ((uint32_t)(FU_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dw = fu_desc_sw32(fu_u32(w_s));
#pragma unroll
        for (int i = 0; i < FU_MT; ++i) {
            const uint64_t da = fu_desc_sw32(fu_u32(a_in + buf * FU_A_BYTES + i * 128 * FU_K * 2));
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
                    tmem_base + (uint32_t)(i * FU_N)),
                "l"(da), "l"(dw), "r"(idesc), "r"(0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fu_u32(&bars->mma_done))
                     : "memory");
    };

    // prologue: weights, the first two input windows, the first tile's MMAs
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fu_u32(&bars->w_full)), "r"(2048u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         fu_u32(w_s)),
                     "l"(&tmap_w), "r"(fu_u32(&bars->w_full)), "r"(0), "r"(0)
                     : "memory");
        if ((int)blockIdx.x < n_tiles) issue_tile(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < n_tiles) issue_tile(blockIdx.x + gridDim.x, 1);
        if ((int)blockIdx.x < n_tiles) {
            fu_wait(&bars->w_full, 0);
            fu_wait(&bars->full[0], 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_mma(0);
        }
    }

    const int quarter = warp & 3, half = warp >> 2;       // TMEM lane quarter / 32-column half of this warp
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const int iy0 = ty * FU_TO * 2 - 1, ix0 = tx * FU_TO * 2 - 1;
        // ---- drain: accumulators -> bias, activation, zero outside the image -> bf16 tile in shared memory ----
        fu_wait(&bars->mma_done, it & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int i = 0; i < FU_MT; ++i) {
            const int m = i * 128 + quarter * 32 + lane;
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * FU_N + half * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < FU_PIX) {
                const int r = (m * 241) >> 12, c = m - r * FU_IT;           // m / 17 for m < 289
                const bool inside = (unsigned)(iy0 + r) < (unsigned)H && (unsigned)(ix0 + c) < (unsigned)W;
                unsigned char* dst = exp_t + m * (FU_N * 2);
                if (!inside) {                     // the depthwise convolution's zero padding
#pragma unroll
                    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dst + (((half * 4 + q) ^ (m & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 bb = *reinterpret_cast<const float2*>(bpw + half * 32 + q * 8 + 2 * e);
                            const float2 tt = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1])),
                                                         make_float2(1.f, 1.f), bb);
                            if constexpr (ACT_PW == DN_ACT_HSWISH) {
                                w[e] = float2_to_h2(fu_act<ACT_PW>(tt.x), fu_act<ACT_PW>(tt.y));
                            } else {               // ReLU / ReLU6 on the packed pair after the (monotone) rounding
                                dn_half2_t h = floats_to_half2(tt.x, tt.y);
                                if constexpr (ACT_PW == DN_ACT_RELU || ACT_PW == DN_ACT_RELU6) h = __hmax2(h, half2_const(0.f));
                                if constexpr (ACT_PW == DN_ACT_RELU6) h = __hmin2(h, half2_const(6.f));
                                w[e] = *reinterpret_cast<uint32_t*>(&h);
                            }
                        }
                        *reinterpret_cast<uint4*>(dst + (((half * 4 + q) ^ (m & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                            // expanded tile complete; TMEM and the input buffer are free again
        if (threadIdx.x == 32) {                    // next tile's MMAs run under this tile's stencil
            const int tn = t + gridDim.x;
            if (tn < n_tiles) {
                fu_wait(&bars->full[buf ^ 1], ((it + 1) >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue_mma(buf ^ 1);
            }
            const int tnn = t + 2 * gridDim.x;      // and the window after that goes into the buffer just consumed
            if (tnn < n_tiles) issue_tile(tnn, buf);
        }
        // ---- stencil: 3x3 stride 2 over the shared-memory tile; thread = 4 channels x a 2 x 2 block of outputs, so the
        // 5 x 5 input window is read and unpacked once for four outputs (25 loads instead of 36) ----
        {
            const int cq = threadIdx.x & 15, blk = threadIdx.x >> 4;        // channel quad 0..15, output block 0..15
            const int by = blk >> 2, bx = blk & 3;                          // 4 x 4 blocks of 2 x 2 outputs
            float2 acc[2][2][2];
            {
                const float4 bq = *reinterpret_cast<const float4*>(bdw + cq * 4);
#pragma unroll
                for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                    for (int b2 = 0; b2 < 2; ++b2) acc[a2][b2][0] = make_float2(bq.x, bq.y), acc[a2][b2][1] = make_float2(bq.z, bq.w);
            }
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                float2 in[5][2];
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int m = (4 * by + r) * FU_IT + 4 * bx + c;
                    const uint2 xv = *reinterpret_cast<const uint2*>(exp_t + m * (FU_N * 2) + (((cq >> 1) ^ (m & 7)) << 4) + (cq & 1) * 8);
                    in[c][0] = h2_to_float2(xv.x), in[c][1] = h2_to_float2(xv.y);
                }
#pragma unroll
                for (int a2 = 0; a2 < 2; ++a2) {                 // output row 2*by + a2 takes input rows 2*a2 .. 2*a2 + 2 of the window
                    const int kh = r - 2 * a2;
                    if (kh < 0 || kh > 2) continue;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 wq = *reinterpret_cast<const float4*>(wdw + (kh * 3 + kw) * FU_N + cq * 4);
                        const float2 w0 = make_float2(wq.x, wq.y), w1 = make_float2(wq.z, wq.w);
#pragma unroll
                        for (int b2 = 0; b2 < 2; ++b2) {
                            acc[a2][b2][0] = __ffma2_rn(in[2 * b2 + kw][0], w0, acc[a2][b2][0]);
                            acc[a2][b2][1] = __ffma2_rn(in[2 * b2 + kw][1], w1, acc[a2][b2][1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    const int gy = ty * FU_TO + 2 * by + a2, gx = tx * FU_TO + 2 * bx + b2;
                    if (gy < Ho && gx < Wo) {
                        uint2 ov;
                        ov.x = float2_to_h2(fu_act<ACT_DW>(acc[a2][b2][0].x), fu_act<ACT_DW>(acc[a2][b2][0].y));
                        ov.y = float2_to_h2(fu_act<ACT_DW>(acc[a2][b2][1].x), fu_act<ACT_DW>(acc[a2][b2][1].y));
                        reinterpret_cast<uint2*>(y)[(((long long)b * Ho + gy) * Wo + gx) * (FU_N / 4) + cq] = ov;
                    }
                }
        }
        __syncthreads();                            // the next drain overwrites the expanded tile
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)FU_TMEM_COLS) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
bool pwdw_fused_supported(int H, int W, int K, int N, int ksize, int stride) {
    return K == FU_K && N == FU_N && ksize == 3 && stride == 2 && H >= 8 && W >= 8;
}

size_t pwdw_fused_smem() { return 2 * FU_A_BYTES + 2048 + ((FU_EXP_BYTES + 127) & ~127) + (9 * FU_N + 2 * FU_N) * 4 + sizeof(FuBars) + 1024; }

int pwdw_fused_make_tmaps(CUtensorMap* tx, CUtensorMap* tw, const void* x, const void* w_pw, int B, int H, int W) {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static PFN fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN)p;
    }
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    DN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, DN_ERR_INVALID, "fused expand input must be 16-byte aligned");
    cuuint64_t gdim[4] = {(cuuint64_t)FU_K, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstride[3] = {(cuuint64_t)FU_K * 2, (cuuint64_t)W * FU_K * 2, (cuuint64_t)H * W * FU_K * 2};
    cuuint32_t box[4] = {(cuuint32_t)FU_K, (cuuint32_t)FU_IT, (cuuint32_t)FU_IT, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tx, DN_TMAP_HALF, 4, const_cast<void*>(x), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled (fused expand input) failed (%d)", (int)r);
    return make_tmap_h16_2d(tw, w_pw, FU_N, FU_K, FU_N, FU_K);
}

template <int ACT_PW, int ACT_DW>
static int fu_launch_t(const CUtensorMap& tx, const CUtensorMap& tw, const float* b_pw, const float* w_dw, const float* b_dw, void* y,
                       int B, int H, int W, cudaStream_t stream) {
    auto kern = pwdw_fused_kernel<ACT_PW, ACT_DW>;
    const size_t smem = pwdw_fused_smem();
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(kern, smem));
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int tiles_x = ceil_div(Wo, FU_TO), tiles_y = ceil_div(Ho, FU_TO);
    const long long n_tiles = (long long)B * tiles_x * tiles_y;
    DN_REQUIRE(n_tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "fused expand + depthwise problem too large");
    long long grid = 2ll * sm_count();
    if (grid > n_tiles) grid = n_tiles;
    launch_pdl(kern, (unsigned)grid, FU_THREADS, smem, stream, tx, tw, b_pw, w_dw, b_dw, (uint4*)y, H, W, Ho, Wo, tiles_x, tiles_y,
               (int)n_tiles);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

int pwdw_fused_launch(const CUtensorMap& tx, const CUtensorMap& tw, const float* b_pw, const float* w_dw, const float* b_dw, void* y,
                      int B, int H, int W, int act_pw, int act_dw, cudaStream_t stream) {
#define DN_FU_CASE(A, D) \
    if (act_pw == A && act_dw == D) return fu_launch_t<A, D>(tx, tw, b_pw, w_dw, b_dw, y, B, H, W, stream)
    DN_FU_CASE(DN_ACT_RELU, DN_ACT_RELU);
    DN_FU_CASE(DN_ACT_RELU6, DN_ACT_RELU6);
    DN_FU_CASE(DN_ACT_HSWISH, DN_ACT_HSWISH);
#undef DN_FU_CASE
    DN_REQUIRE(false, DN_ERR_UNSUPPORTED, "fused expand + depthwise: unsupported activation pair (%d, %d)", act_pw, act_dw);
}

}  // namespace dn

using namespace dn;

extern "C" int dn_pwdw_fused(const void* x, const void* w_pw, const float* b_pw, const float* w_dw, const float* b_dw, void* y,
                             int B, int H, int W, int K, int N, int ksize, int stride, int act_pw, int act_dw, void* stream_) {
    DN_REQUIRE(x && w_pw && b_pw && w_dw && b_dw && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(pwdw_fused_supported(H, W, K, N, ksize, stride), DN_ERR_UNSUPPORTED,
               "fused expand + depthwise supports K = 16, N = 64, 3x3 stride 2 (got K=%d N=%d k=%d s=%d)", K, N, ksize, stride);
    CUtensorMap tx, tw;
    int rc = pwdw_fused_make_tmaps(&tx, &tw, x, w_pw, B, H, W);
    if (rc) return rc;
    return pwdw_fused_launch(tx, tw, b_pw, w_dw, b_dw, y, B, H, W, act_pw, act_dw, (cudaStream_t)stream_);
}
