// Fused depthwise 3x3 stride-1 (+ BN + act) -> pointwise project (1x1 conv + BN, no activation) (+ residual) for
// sm_100a: the depthwise output never leaves the SM.  The other half of DESIGN.md section 8 item 1, for MobileNetV3
// block 1 (16 channels at 160x160: dw 3x3 + ReLU, 16 -> 16 project, `result += input`): unfused the block reads the
// input twice (depthwise, residual) and writes + re-reads the depthwise output -- 1.05 GB per batch of 256 -- fused it
// reads the input once and writes the output once (0.42 GB).
// Reference: InvertedResidual.forward, demonet/models/mobilenetv3.py:80-99 (see dn_dwconv / dn_pwconv).
//
// A persistent CTA walks 16 x 32 output tiles (512 pixels = four MMA tiles; the fixed latencies of a tile -- barrier
// round trips, MMA commit, TMEM load -- are paid once per 512 pixels).  Per tile:
//   TMA      one cp.async.bulk.tensor.4d fetches the 10 x 18 x 16-channel input window (zero fill = the depthwise
//            padding), double-buffered
//   stencil  thread = 4 channels x 4 vertically adjacent pixels: the 6 x 3 window is read once, taps in registers, packed
//            FFMA2, bias + activation, bf16, written as 8-byte pieces of the K-major SWIZZLE_32B A operand (row = pixel)
//   MMA      four tcgen05.mma (128 x 16 x 16, kind::f16) -> fp32 accumulators in TMEM (4 x 16 columns)
//   drain    tcgen05.ld -> + bias (+ residual = the window's centre pixel, already in shared memory) -> bf16 -> global
#include <cuda.h>

#include "common.cuh"
#include "dwconv.cuh"
#include "pwconv.cuh"

namespace dn {

constexpr int FD_TH = 16, FD_TW = 32;                   // output tile
constexpr int FD_SUB = FD_TH * FD_TW / 128;             // MMA tiles (128 pixels each) per output tile
constexpr int FD_IH = FD_TH + 2, FD_IW = FD_TW + 2;     // input window (k = 3, stride 1)
constexpr int FD_C = 16;
constexpr int FD_THREADS = 256;
constexpr int FD_IN_BYTES = ((FD_IH * FD_IW * FD_C * 2) + 127) & ~127;     // 5760 -> 5760 (multiple of 128)
constexpr int FD_A_BYTES = FD_SUB * 128 * FD_C * 2;     // 16 KiB
constexpr int FD_TMEM_COLS = 64;                        // FD_SUB x 16 accumulator columns

__device__ __forceinline__ uint32_t fd_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fd_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(fd_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
template <int ACT>
__device__ __forceinline__ float fd_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));
    return v;
}
__device__ __forceinline__ uint64_t fd_desc_sw32(uint32_t smem_addr) {        // K-major SWIZZLE_32B, 8-row groups 256 B apart
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}

struct __align__(8) FdBars {
    uint64_t full[2];          // TMA -> stencil: input window landed
    uint64_t mma_done;         // MMA -> drain
    uint64_t w_full;           // project weights landed
    uint32_t tmem_base;
    uint32_t pad;
};

template <int ACT_DW, bool RESIDUAL>
__global__ void __launch_bounds__(FD_THREADS, 3)
dwpw_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const float* __restrict__ w_dw, const float* __restrict__ b_dw, const float* __restrict__ b_pw,
                  uint4* __restrict__ y, int H, int W, int tiles_x, int tiles_y, int n_tiles) {
    extern __shared__ __align__(1024) unsigned char fd_smem_raw[];
    // aligned by pointer arithmetic on the __shared__ array (not through an integer cast) so that the compiler keeps the
    // shared address space and emits LDS / STS instead of generic loads and stores
    unsigned char* base = fd_smem_raw + ((1024u - (fd_u32(fd_smem_raw) & 1023u)) & 1023u);
    unsigned char* a_t = base;                                       // [FD_SUB][128 x 16] bf16, SWIZZLE_32B (MMA A operands)
    unsigned char* w_s = a_t + FD_A_BYTES;                           // [16 x 16] bf16, SWIZZLE_32B (MMA B operand)
    unsigned char* in_t = w_s + 1024;                                // [2][10 x 18 x 16] bf16 input windows
    float* wdw = reinterpret_cast<float*>(in_t + 2 * FD_IN_BYTES);   // [9][16]
    float* bdw = wdw + 9 * FD_C;                                     // [16]
    float* bpw = bdw + FD_C;                                         // [16]
    FdBars* bars = reinterpret_cast<FdBars*>(bpw + FD_C);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fd_u32(&bars->full[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fd_u32(&bars->full[1])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fd_u32(&bars->mma_done)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fd_u32(&bars->w_full)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fd_u32(&bars->tmem_base)),
                     "r"((uint32_t)FD_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 9 * FD_C; i += FD_THREADS) wdw[i] = __ldg(w_dw + i);
    if (threadIdx.x < FD_C) bdw[threadIdx.x] = __ldg(b_dw + threadIdx.x), bpw[threadIdx.x] = __ldg(b_pw + threadIdx.x);
    pdl_trigger();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    auto issue_tile = [&](int t, int buf) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const uint32_t bar = fd_u32(&bars->full[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(FD_IH * FD_IW * FD_C * 2))
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                "r"(fd_u32(in_t + buf * FD_IN_BYTES)),
            "l"(&tmap_x), "r"(bar), "r"(0), "r"(tx * FD_TW - 1), "r"(ty * FD_TH - 1), "r"(b)
            : "memory");
    };
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fd_u32(&bars->w_full)), "r"(512u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         fd_u32(w_s)),
                     "l"(&tmap_w), "r"(fd_u32(&bars->w_full)), "r"(0), "r"(0)
                     : "memory");
        if ((int)blockIdx.x < n_tiles) issue_tile(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < n_tiles) issue_tile(blockIdx.x + gridDim.x, 1);
    }

    // drain role: TMEM lane quarter and column half
    const int quarter = warp & 3, colh = warp >> 2;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const unsigned char* win = in_t + buf * FD_IN_BYTES;
        fd_wait(&bars->full[buf], (it >> 1) & 1u);
        // ---- stencil -> A operands.  Thread = 4 channels x 4 vertically adjacent pixels of one column: the 6 x 3 window is
        // read once (18 8-byte loads for 4 outputs), the 9 x 4 taps live in registers, lanes walk the channels of a pixel
        // and then the columns, so shared-memory reads are conflict-free ----
        {
            const int cq = threadIdx.x & 3, col = (threadIdx.x >> 2) & 31, rg = threadIdx.x >> 7;
            float2 wv[9][2];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float4 w4 = *reinterpret_cast<const float4*>(wdw + tap * FD_C + cq * 4);
                wv[tap][0] = make_float2(w4.x, w4.y), wv[tap][1] = make_float2(w4.z, w4.w);
            }
            const float4 bq = *reinterpret_cast<const float4*>(bdw + cq * 4);
#pragma unroll 1
            for (int rb = rg; rb < FD_TH / 4; rb += 2) {
                const int r0 = rb * 4;
                float2 acc[4][2];
#pragma unroll
                for (int o = 0; o < 4; ++o) acc[o][0] = make_float2(bq.x, bq.y), acc[o][1] = make_float2(bq.z, bq.w);
#pragma unroll
                for (int wr = 0; wr < 6; ++wr) {
                    float2 in[3][2];
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const uint2 xv = *reinterpret_cast<const uint2*>(win + ((r0 + wr) * FD_IW + col + kw) * (FD_C * 2) + cq * 8);
                        in[kw][0] = h2_to_float2(xv.x), in[kw][1] = h2_to_float2(xv.y);
                    }
#pragma unroll
                    for (int o = 0; o < 4; ++o) {                    // output row r0 + o takes window row wr with kh = wr - o
                        const int kh = wr - o;
                        if (kh < 0 || kh > 2) continue;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            acc[o][0] = __ffma2_rn(in[kw][0], wv[kh * 3 + kw][0], acc[o][0]);
                            acc[o][1] = __ffma2_rn(in[kw][1], wv[kh * 3 + kw][1], acc[o][1]);
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int m = (r0 + o) * FD_TW + col;
                    uint2 av;
                    av.x = float2_to_h2(fd_act<ACT_DW>(acc[o][0].x), fd_act<ACT_DW>(acc[o][0].y));
                    av.y = float2_to_h2(fd_act<ACT_DW>(acc[o][1].x), fd_act<ACT_DW>(acc[o][1].y));
                    // SWIZZLE_32B: 16-byte chunk c of row m lives at chunk position c ^ ((m >> 2) & 1)
                    *reinterpret_cast<uint2*>(a_t + m * 32 + (((cq >> 1) ^ ((m >> 2) & 1)) << 4) + (cq & 1) * 8) = av;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the MMA
        __syncthreads();
        if (threadIdx.x == 32) {
            if (it == 0) fd_wait(&bars->w_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // instruction descriptor: D = f32, A = B = bf16, K-major both, N = 16, M = 128
            constexpr uint32_t idesc = (1u << 4) | (DN_UMMA_AB_FORMAT << 7) | (DN_UMMA_AB_FORMAT << 10) | ((uint32_t)(FD_C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t dw = fd_desc_sw32(fd_u32(w_s));
#pragma unroll
            for (int j = 0; j < FD_SUB; ++j)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
                        tmem_base + (uint32_t)(j * FD_C)),
                    "l"(fd_desc_sw32(fd_u32(a_t + j * 128 * FD_C * 2))), "l"(dw), "r"(idesc), "r"(0u)
                    : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fd_u32(&bars->mma_done))
                         : "memory");
        }
        // ---- drain: accumulators -> + bias (+ residual) -> bf16 -> global ----
        fd_wait(&bars->mma_done, it & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            uint32_t v[FD_SUB][8];
#pragma unroll
            for (int j = 0; j < FD_SUB; ++j) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * FD_C + colh * 8);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(v[j][0]), "=r"(v[j][1]), "=r"(v[j][2]), "=r"(v[j][3]), "=r"(v[j][4]), "=r"(v[j][5]), "=r"(v[j][6]),
                               "=r"(v[j][7])
                             : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < FD_SUB; ++j) {
                const int m = j * 128 + quarter * 32 + lane;             // pixel of the tile = TMEM lane of MMA tile j
                const int r = m / FD_TW, c = m % FD_TW;
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j][e]) + bpw[colh * 8 + e];
                if (RESIDUAL) {                                          // `result += input`: the window's centre pixel
                    float rsd[8];
                    unpack8(*reinterpret_cast<const uint4*>(win + ((r + 1) * FD_IW + c + 1) * (FD_C * 2) + colh * 16), rsd);
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] += rsd[e];
                }
                const int gy = ty * FD_TH + r, gx = tx * FD_TW + c;
                if (gy < H && gx < W) y[(((long long)b * H + gy) * W + gx) * 2 + colh] = pack8(f);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                            // window, A tile and accumulator are free again
        if (threadIdx.x == 32) {
            const int tnn = t + 2 * gridDim.x;
            if (tnn < n_tiles) issue_tile(tnn, buf);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)FD_TMEM_COLS) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
bool dwpw_fused_supported(int H, int W, int C, int N, int ksize, int stride) {
    return C == FD_C && N == FD_C && ksize == 3 && stride == 1 && H >= 8 && W >= 16;      // smaller maps: not worth a tile
}

static size_t dwpw_fused_smem() { return FD_A_BYTES + 1024 + 2 * FD_IN_BYTES + (9 * FD_C + 2 * FD_C) * 4 + sizeof(FdBars) + 1024; }

int dwpw_fused_make_tmaps(CUtensorMap* tx, CUtensorMap* tw, const void* x, const void* w_pw, int B, int H, int W) {
    DwTiling tl{};
    tl.CB = FD_C, tl.IWT = FD_IW, tl.IHT = FD_IH;
    int rc = dw_make_tmap(tx, x, B, H, W, FD_C, tl);            // plain NHWC window, zero fill outside the image
    if (rc) return rc;
    return make_tmap_h16_2d(tw, w_pw, FD_C, FD_C, FD_C, FD_C);  // [N = 16, K = 16], SWIZZLE_32B
}

template <int ACT_DW, bool RESIDUAL>
static int fd_launch_t(const CUtensorMap& tx, const CUtensorMap& tw, const float* w_dw, const float* b_dw, const float* b_pw, void* y,
                       int B, int H, int W, cudaStream_t stream) {
    auto kern = dwpw_fused_kernel<ACT_DW, RESIDUAL>;
    const size_t smem = dwpw_fused_smem();
    static SmemOptIn optin;
    int per_sm = 1;
    DN_CHECK_CUDA(optin.ensure(kern, smem));
    DN_CHECK_CUDA(optin.blocks_per_sm_tmem(kern, FD_THREADS, smem, FD_TMEM_COLS, &per_sm));
    const int tiles_x = ceil_div(W, FD_TW), tiles_y = ceil_div(H, FD_TH);
    const long long n_tiles = (long long)B * tiles_x * tiles_y;
    DN_REQUIRE(n_tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "fused depthwise + project problem too large");
    long long grid = (long long)per_sm * sm_count();
    if (grid > n_tiles) grid = n_tiles;
    launch_pdl(kern, (unsigned)grid, FD_THREADS, smem, stream, tx, tw, w_dw, b_dw, b_pw, (uint4*)y, H, W, tiles_x, tiles_y, (int)n_tiles);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

int dwpw_fused_launch(const CUtensorMap& tx, const CUtensorMap& tw, const float* w_dw, const float* b_dw, const float* b_pw, void* y,
                      int B, int H, int W, int act_dw, int residual, cudaStream_t stream) {
#define DN_FD_CASE(A)                                                                                             \
    if (act_dw == A)                                                                                              \
        return residual ? fd_launch_t<A, true>(tx, tw, w_dw, b_dw, b_pw, y, B, H, W, stream)                      \
                        : fd_launch_t<A, false>(tx, tw, w_dw, b_dw, b_pw, y, B, H, W, stream)
    DN_FD_CASE(DN_ACT_RELU);
    DN_FD_CASE(DN_ACT_RELU6);
    DN_FD_CASE(DN_ACT_HSWISH);
#undef DN_FD_CASE
    DN_REQUIRE(false, DN_ERR_UNSUPPORTED, "fused depthwise + project: unsupported depthwise activation %d", act_dw);
}

}  // namespace dn

using namespace dn;

extern "C" int dn_dwpw_fused(const void* x, const float* w_dw, const float* b_dw, const void* w_pw, const float* b_pw, void* y,
                             int B, int H, int W, int C, int N, int ksize, int stride, int act_dw, int residual, void* stream_) {
    DN_REQUIRE(x && w_dw && b_dw && w_pw && b_pw && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0, DN_ERR_INVALID, "bad shape");
    DN_REQUIRE(x != y, DN_ERR_INVALID, "fused depthwise + project cannot run in place");
    DN_REQUIRE(dwpw_fused_supported(H, W, C, N, ksize, stride), DN_ERR_UNSUPPORTED,
               "fused depthwise + project supports C = N = 16, 3x3 stride 1 (got C=%d N=%d k=%d s=%d)", C, N, ksize, stride);
    CUtensorMap tx, tw;
    int rc = dwpw_fused_make_tmaps(&tx, &tw, x, w_pw, B, H, W);
    if (rc) return rc;
    return dwpw_fused_launch(tx, tw, w_dw, b_dw, b_pw, y, B, H, W, act_dw, residual, (cudaStream_t)stream_);
}
