// Dense 3 x 3 convolution (+ bias + activation) as an implicit GEMM on the tcgen05 tensor cores, sm_100a.
// the kernels ssd300_vgg16 needs beyond the SSDLite path -- VGG16 features and the SSD extra blocks
// (demonet/models/ssd_vgg16.py:30-109; stride 1 / 2, padding 0 / 1, conv6 with dilation 6) and the dense SSD heads
// (SSDScoringHead with 3x3 convolutions, demonet/models/generalized_ssd.py:38-92).
//
//   D[pixels, N] = sum over the 9 taps (dy, dx) of  A_(dy,dx)[pixels, C] . W_(dy,dx)[N, C]^T
//
// No im2col buffer: the M tile is a patch of TB images x TH x TW OUTPUT pixels (<= 128), and the A operand of tap
// (dy, dx) is ONE 4-D TMA box of the NHWC input, [TB][TH][TW][64 channels], fetched at the patch origin shifted by
// (dy, dx) * dilation - padding (traversal stride = the convolution's stride).  Pixels outside the image are the tensor
// map's out-of-bounds zero fill = the convolution's zero padding, so there is not one boundary branch in the kernel.
// The box lands in shared memory as [rows = pixels][64 channels = 128 bytes], SWIZZLE_128B: exactly the K-major A tile
// the pointwise GEMM uses, and from there on the kernel IS that GEMM with a 9 x C / 64 long K loop: warp 0 TMA producer,
// warp 1 MMA issuer (UMMA 128 x BLOCK_N x 16, fp32 accumulators in TMEM, two buffers), warps 2..9 epilogue (thread = patch
// pixel; bias + activation, staged through shared memory so that stores are row-contiguous; fp16 / bf16 NHWC output or the
// head's strided fp32 [B, P, K] layout).  Rows of a patch that fall outside the image or the batch are computed on whatever
// the box holds and never stored.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "pwconv.cuh"
#include "tcgen05.cuh"

namespace dn {

constexpr int C3_BLOCK_M = 128;
constexpr int C3_BLOCK_K = 64;
constexpr int C3_MAX_STAGES = 4;
constexpr int C3_EPI_WARPS = 8;
constexpr int C3_THREADS = 64 + 32 * C3_EPI_WARPS;
constexpr int C3_A_STAGE_BYTES = C3_BLOCK_M * C3_BLOCK_K * 2;
constexpr int C3_STAGE_PITCH = 36;
constexpr int C3_STAGE_BYTES = 5120;
constexpr int C3_STATIC_SMEM = C3_EPI_WARPS * C3_STAGE_BYTES + 1024;

struct ConvGeom {
    int B, H, W, C;              // input  [B, H, W, C]
    int Ho, Wo, N;               // output [B, Ho, Wo, N]
    int stride, pad, dil;
    int TB, TH, TW;              // M tile: TB images x TH x TW output pixels (TB * TH * TW <= 128)
    int tiles_x, tiles_y, tiles_b;
};

struct __align__(8) C3Barriers {
    uint64_t full[C3_MAX_STAGES];
    uint64_t empty[C3_MAX_STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int ACT>
__device__ __forceinline__ float c3_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}


// the storing half of the epilogue: a warp's 32 rows x 32 staged columns, lanes = 8 column groups x 4 rows per pass, so that a
// row's 32 columns leave as one contiguous run; `out_row` is each lane's own row offset (-1: not stored), shared by shuffle
__device__ __forceinline__ void c3_store_rows(const PwEpilogue& ep, const float* stg, long long out_row, int lane, int col0, int col_end) {
    const int cg = (lane & 7) * 4;
    const int n = col0 + cg;
    const int cnt = col_end - n;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3);
        const long long orow = __shfl_sync(0xffffffffu, out_row, r);
        if (orow >= 0 && cnt > 0) {
            const float4 o = *reinterpret_cast<const float4*>(stg + r * C3_STAGE_PITCH + cg);
            const float f[4] = {o.x, o.y, o.z, o.w};
            if (ep.out_fp32) {
                float* dst = reinterpret_cast<float*>(ep.y) + orow + n;
                if (cnt >= 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    *reinterpret_cast<float4*>(dst) = o;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < cnt) dst[i] = f[i];
                }
            } else {
                dn_half_t* dst = reinterpret_cast<dn_half_t*>(ep.y) + orow + n;
                if (cnt >= 4 && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
                    uint2 pk;
                    pk.x = float2_to_h2(f[0], f[1]);
                    pk.y = float2_to_h2(f[2], f[3]);
                    *reinterpret_cast<uint2*>(dst) = pk;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < cnt) dst[i] = float_to_half(f[i]);
                }
            }
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(C3_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, PwEpilogue ep, ConvGeom g,
                  int block_n, int n_tiles, int num_tiles, int num_stages, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t c3_smem_raw[];
    __shared__ __align__(1024) uint8_t s_stage_raw[C3_EPI_WARPS][C3_STAGE_BYTES];
    uint8_t* smem = c3_smem_raw + ((1024u - (smem_u32(c3_smem_raw) & 1023u)) & 1023u);
    const int w_stage_bytes = block_n * C3_BLOCK_K * 2;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + num_stages * C3_A_STAGE_BYTES;
    C3Barriers* bars = reinterpret_cast<C3Barriers*>(smem_w + num_stages * w_stage_bytes);
    float* s_bias = reinterpret_cast<float*>(bars + 1);                       // [n_tiles * block_n + 32], zero padded
    for (int i = threadIdx.x; i < n_tiles * block_n + 32; i += blockDim.x) s_bias[i] = (i < g.N) ? __ldg(ep.bias + i) : 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = g.C / C3_BLOCK_K;
    const int k_iters = 9 * kblocks;
    const int rows_used = g.TB * g.TH * g.TW;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->tmem_full[b], 1);
            mbar_init(&bars->tmem_empty[b], C3_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, (uint32_t)tmem_cols);
    pdl_trigger();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    const int tiles_xy = g.tiles_x * g.tiles_y;
    if (warp == 0) {
        // ===== TMA producer: 9 taps x C / 64 k-blocks per tile =====
        if (elect_one()) {
            const uint32_t stage_bytes = (uint32_t)(rows_used * C3_BLOCK_K * 2 + w_stage_bytes);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / n_tiles, n0 = (tile % n_tiles) * block_n;
                const int bt = mt / tiles_xy, rem = mt - bt * tiles_xy;
                const int y0 = (rem / g.tiles_x) * g.TH, x0 = (rem % g.tiles_x) * g.TW, b0 = bt * g.TB;
                for (int tap = 0; tap < 9; ++tap) {
                    const int iy = y0 * g.stride + (tap / 3) * g.dil - g.pad;
                    const int ix = x0 * g.stride + (tap % 3) * g.dil - g.pad;
                    for (int kb = 0; kb < kblocks; ++kb, ++it) {
                        const int s = it % num_stages;
                        mbar_wait_producer(&bars->empty[s], ((it / num_stages) & 1u) ^ 1u);
                        mbar_expect_tx(&bars->full[s], stage_bytes);
                        tma_load_4d(smem_a + s * C3_A_STAGE_BYTES, &tmap_a, &bars->full[s], kb * C3_BLOCK_K, ix, iy, b0);
                        tma_load_2d(smem_w + s * w_stage_bytes, &tmap_w, &bars->full[s], kb * C3_BLOCK_K, tap * g.N + n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(C3_BLOCK_M, block_n);
        uint32_t it = 0, lt = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            mbar_wait(&bars->tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * (uint32_t)block_n;
            for (int ki = 0; ki < k_iters; ++ki, ++it) {
                const int s = it % num_stages;
                mbar_wait(&bars->full[s], (it / num_stages) & 1u);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_smem_desc(smem_u32(smem_a + s * C3_A_STAGE_BYTES));
                    const uint64_t dw = make_smem_desc(smem_u32(smem_w + s * w_stage_bytes));
#pragma unroll
                    for (int k = 0; k < C3_BLOCK_K / 16; ++k)
                        umma_f16(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (ki | k) != 0 ? 1u : 0u);
                    umma_commit(&bars->empty[s]);
                    if (ki == k_iters - 1) umma_commit(&bars->tmem_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: thread = patch pixel (TMEM lane) =====
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int thw = g.TH * g.TW;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            const int mt = tile / n_tiles, n0 = (tile % n_tiles) * block_n;
            const int bt = mt / tiles_xy, rem = mt - bt * tiles_xy;
            const int y0 = (rem / g.tiles_x) * g.TH, x0 = (rem % g.tiles_x) * g.TW, b0 = bt * g.TB;
            // this lane's own output pixel; shared with the storing lanes by shuffle below (-1 = nothing to store)
            long long out_row = -1;
            {
                const int tb = row / thw, rr = row - tb * thw;
                const int y = y0 + rr / g.TW, x = x0 + rr % g.TW, b = b0 + tb;
                if (row < rows_used && b < g.B && y < g.Ho && x < g.Wo) out_row = ep.row_offset((b * g.Ho + y) * g.Wo + x);
            }
            const int n_valid = min(block_n, g.N - n0);
            mbar_wait(&bars->tmem_full[buf], (lt >> 1) & 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * (uint32_t)block_n + ((uint32_t)(quarter * 32) << 16);
            float* stg = reinterpret_cast<float*>(s_stage_raw[warp - 2]);
            for (int c0 = half * 32; c0 < n_valid; c0 += 64) {
                uint32_t v[32];
                const bool second = (c0 + 16 < block_n);
                tmem_ld16(tmem_d + (uint32_t)c0, v);
                if (second) tmem_ld16(tmem_d + (uint32_t)(c0 + 16), v + 16);
                const float* sbw = s_bias + n0 + c0;
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = c3_act<ACT>(__uint_as_float(v[j + 0]) + sbw[j + 0]);
                    o.y = c3_act<ACT>(__uint_as_float(v[j + 1]) + sbw[j + 1]);
                    o.z = c3_act<ACT>(__uint_as_float(v[j + 2]) + sbw[j + 2]);
                    o.w = c3_act<ACT>(__uint_as_float(v[j + 3]) + sbw[j + 3]);
                    *reinterpret_cast<float4*>(stg + lane * C3_STAGE_PITCH + j) = o;
                }
                __syncwarp();
                c3_store_rows(ep, stg, out_row, lane, n0 + c0, n0 + n_valid);
                __syncwarp();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars->tmem_empty[buf])) : "memory");
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
    }
}


// ---- halo variant (stride 1, dilation 1, padding 1) ---------------------------------------------------------------
// The 9 taps of a patch read the SAME (TH + 2) x (TW + 2) input window, shifted.  The kernel above fetches that window nine
// times (nine boxes of 128 rows from L2 per k-block: at 64 / 128 channels the layer is bound by L2 -> shared-memory traffic,
// not by the tensor pipe).  Here the window is fetched ONCE per k-block as a [TH + 2][PW = TW + 2][64] box, and the M tile
// is laid out with the window's own row pitch: GEMM row i <-> patch pixel (i / PW, i % PW); the PW - TW = 2 rightmost
// columns of every patch row are dummies that are computed and never stored.  With that pitch the A operand of tap
// (dy, dx) is the SAME shared-memory tile started (dy * PW + dx) rows later -- a descriptor start-address offset of
// (dy * PW + dx) * 128 bytes -- so one 24 KB box feeds 36 UMMAs instead of 4.  (Measured on B200: the 128-byte swizzle is
// a function of the ABSOLUTE shared-memory address bits [7,10), so a start address that is not a multiple of 8 rows needs
// nothing else; setting the descriptor's base-offset field to (address >> 7) & 7 on top of that gives wrong products.)  The filter is either resident for the whole kernel (9 * C * N * 2 bytes fit next to the window
// ring: conv1_2) or streamed through its own ring by a second producer warp.
constexpr int CH_THREADS = C3_THREADS + 32;          // + warp 10: filter producer
constexpr int CH_MAX_A = 4, CH_MAX_W = 8;

struct HaloGeom {
    int B, H, W, C, N;
    int TH, TW, PW;              // patch, row pitch PW = TW + 2
    int tiles_x, tiles_y;
    int a_stage_bytes;           // (128 + 2 * PW + 2) rows of 128 B, rounded up to 1024
    int w_resident;              // 1: all 9 * C / 64 filter tiles stay in shared memory (n_tiles == 1)
};

struct __align__(8) HaloBarriers {
    uint64_t a_full[CH_MAX_A], a_empty[CH_MAX_A];
    uint64_t w_full[CH_MAX_W], w_empty[CH_MAX_W];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base, pad;
};

template <int ACT>
__global__ void __launch_bounds__(CH_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, PwEpilogue ep, HaloGeom g,
                    int block_n, int n_tiles, int num_tiles, int na, int nw, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t c3_smem_raw[];
    __shared__ __align__(1024) uint8_t s_stage_raw[C3_EPI_WARPS][C3_STAGE_BYTES];
    uint8_t* smem = c3_smem_raw + ((1024u - (smem_u32(c3_smem_raw) & 1023u)) & 1023u);
    const int kblocks = g.C / C3_BLOCK_K;
    const int w_tile_bytes = block_n * C3_BLOCK_K * 2;
    const int w_slots = g.w_resident ? 9 * kblocks : nw;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + na * g.a_stage_bytes;
    HaloBarriers* bars = reinterpret_cast<HaloBarriers*>(smem_w + (size_t)w_slots * w_tile_bytes);
    float* s_bias = reinterpret_cast<float*>(bars + 1);
    for (int i = threadIdx.x; i < n_tiles * block_n + 32; i += blockDim.x) s_bias[i] = (i < g.N) ? __ldg(ep.bias + i) : 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < CH_MAX_A; ++s) {
            mbar_init(&bars->a_full[s], 1);
            mbar_init(&bars->a_empty[s], 1);
        }
        for (int s = 0; s < CH_MAX_W; ++s) {
            mbar_init(&bars->w_full[s], 1);
            mbar_init(&bars->w_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->tmem_full[b], 1);
            mbar_init(&bars->tmem_empty[b], C3_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, (uint32_t)tmem_cols);
    pdl_trigger();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    const int tiles_xy = g.tiles_x * g.tiles_y;
    if (warp == 0) {
        // ===== window producer: one box per (tile, k-block) =====
        if (elect_one()) {
            const uint32_t box_bytes = (uint32_t)((g.TH + 2) * g.PW * C3_BLOCK_K * 2);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / n_tiles;
                const int b = mt / tiles_xy, rem = mt - b * tiles_xy;
                const int y0 = (rem / g.tiles_x) * g.TH, x0 = (rem % g.tiles_x) * g.TW;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % na;
                    mbar_wait_producer(&bars->a_empty[s], ((it / na) & 1u) ^ 1u);
                    mbar_expect_tx(&bars->a_full[s], box_bytes);
                    tma_load_4d(smem_a + (size_t)s * g.a_stage_bytes, &tmap_a, &bars->a_full[s], kb * C3_BLOCK_K, x0 - 1, y0 - 1, b);
                }
            }
        }
    } else if (warp == 10) {
        // ===== filter producer: resident (loaded once) or one tile per (tile, k-block, tap) =====
        if (elect_one()) {
            if (g.w_resident) {
                mbar_expect_tx(&bars->w_full[0], (uint32_t)(9 * kblocks * w_tile_bytes));
                for (int kb = 0; kb < kblocks; ++kb)
                    for (int tap = 0; tap < 9; ++tap)
                        tma_load_2d(smem_w + (size_t)(kb * 9 + tap) * w_tile_bytes, &tmap_w, &bars->w_full[0], kb * C3_BLOCK_K, tap * g.N);
            } else {
                uint32_t it = 0;
                for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                    const int n0 = (tile % n_tiles) * block_n;
                    for (int kb = 0; kb < kblocks; ++kb)
                        for (int tap = 0; tap < 9; ++tap, ++it) {
                            const int s = it % nw;
                            mbar_wait_producer(&bars->w_empty[s], ((it / nw) & 1u) ^ 1u);
                            mbar_expect_tx(&bars->w_full[s], (uint32_t)w_tile_bytes);
                            tma_load_2d(smem_w + (size_t)s * w_tile_bytes, &tmap_w, &bars->w_full[s], kb * C3_BLOCK_K, tap * g.N + n0);
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: per window, 9 taps x 4 UMMAs =====
        const uint32_t idesc = make_idesc(C3_BLOCK_M, block_n);
        uint32_t ita = 0, itw = 0, lt = 0;
        if (g.w_resident) mbar_wait(&bars->w_full[0], 0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            mbar_wait(&bars->tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * (uint32_t)block_n;
            for (int kb = 0; kb < kblocks; ++kb, ++ita) {
                const int sa = ita % na;
                mbar_wait(&bars->a_full[sa], (ita / na) & 1u);
                const uint32_t a_base = smem_u32(smem_a + (size_t)sa * g.a_stage_bytes);
                for (int tap = 0; tap < 9; ++tap) {
                    int sw = kb * 9 + tap;
                    if (!g.w_resident) {
                        sw = itw % nw;
                        mbar_wait(&bars->w_full[sw], (itw / nw) & 1u);
                        ++itw;
                    }
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint64_t da = make_smem_desc(a_base + (uint32_t)(((tap / 3) * g.PW + tap % 3) * 128));
                        const uint64_t dw = make_smem_desc(smem_u32(smem_w + (size_t)sw * w_tile_bytes));
#pragma unroll
                        for (int k = 0; k < C3_BLOCK_K / 16; ++k)
                            umma_f16(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | tap | k) != 0 ? 1u : 0u);
                        if (!g.w_resident) umma_commit(&bars->w_empty[sw]);
                        if (tap == 8) {
                            umma_commit(&bars->a_empty[sa]);
                            if (kb == kblocks - 1) umma_commit(&bars->tmem_full[buf]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue: thread = GEMM row = patch pixel (row / PW, row % PW); dummy columns and rows are not stored =====
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1u;
            const int mt = tile / n_tiles, n0 = (tile % n_tiles) * block_n;
            const int b = mt / tiles_xy, rem = mt - b * tiles_xy;
            const int y0 = (rem / g.tiles_x) * g.TH, x0 = (rem % g.tiles_x) * g.TW;
            long long out_row = -1;
            {
                const int ty = row / g.PW, tx = row - ty * g.PW;
                const int y = y0 + ty, x = x0 + tx;
                if (ty < g.TH && tx < g.TW && y < g.H && x < g.W) out_row = ep.row_offset((b * g.H + y) * g.W + x);
            }
            const int n_valid = min(block_n, g.N - n0);
            mbar_wait(&bars->tmem_full[buf], (lt >> 1) & 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * (uint32_t)block_n + ((uint32_t)(quarter * 32) << 16);
            float* stg = reinterpret_cast<float*>(s_stage_raw[warp - 2]);
            for (int c0 = half * 32; c0 < n_valid; c0 += 64) {
                uint32_t v[32];
                const bool second = (c0 + 16 < block_n);
                tmem_ld16(tmem_d + (uint32_t)c0, v);
                if (second) tmem_ld16(tmem_d + (uint32_t)(c0 + 16), v + 16);
                const float* sbw = s_bias + n0 + c0;
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = c3_act<ACT>(__uint_as_float(v[j + 0]) + sbw[j + 0]);
                    o.y = c3_act<ACT>(__uint_as_float(v[j + 1]) + sbw[j + 1]);
                    o.z = c3_act<ACT>(__uint_as_float(v[j + 2]) + sbw[j + 2]);
                    o.w = c3_act<ACT>(__uint_as_float(v[j + 3]) + sbw[j + 3]);
                    *reinterpret_cast<float4*>(stg + lane * C3_STAGE_PITCH + j) = o;
                }
                __syncwarp();
                c3_store_rows(ep, stg, out_row, lane, n0 + c0, n0 + n_valid);
                __syncwarp();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars->tmem_empty[buf])) : "memory");
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled3 c3_encode_fn() {
    static PFN_encodeTiled3 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled3)p;
    }
    return fn;
}

// Patch shape: TW columns x TH rows (x TB whole images when one image fits) with TB * TH * TW <= 128, chosen to waste as few
// of the 128 MMA rows as possible over the whole layer.
static void c3_pick_patch(int B, int Ho, int Wo, int* tb_out, int* th_out, int* tw_out) {
    double best = -1.0;
    int btb = 1, bth = 1, btw = 1;
    for (int tw = 1; tw <= std::min(Wo, 128); ++tw) {
        const int th = std::min(Ho, 128 / tw);
        if (th < 1) break;
        int tb = 1;
        if (tw == Wo && th == Ho) tb = std::max(1, std::min(B, 128 / (tw * th)));
        const long long tiles = (long long)ceil_div(Wo, tw) * ceil_div(Ho, th) * ceil_div(B, tb);
        const double util = (double)B * Ho * Wo / ((double)tiles * 128.0);
        if (util > best + 1e-9) best = util, btb = tb, bth = th, btw = tw;
    }
    *tb_out = btb, *th_out = bth, *tw_out = btw;
}

template <int ACT>
static int c3_launch(const CUtensorMap& ta, const CUtensorMap& tw, const PwEpilogue& ep, const ConvGeom& g, int bn, int nt,
                     int tiles, int st, int cols, unsigned grid, size_t smem, cudaStream_t stream) {
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(conv3x3_tc_kernel<ACT>, smem));
    launch_pdl(conv3x3_tc_kernel<ACT>, grid, C3_THREADS, smem, stream, ta, tw, ep, g, bn, nt, tiles, st, cols);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

// ---- halo variant, host side ----
static int c3_halo_mode() {                      // DN_C3_HALO: 0 off, 1 on (default)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DN_C3_HALO");
        v = e ? atoi(e) : 1;
    }
    return v;
}

template <int ACT>
static int ch_launch(const CUtensorMap& ta, const CUtensorMap& tw, const PwEpilogue& ep, const HaloGeom& g, int bn, int nt, int tiles,
                     int na, int nw, int cols, unsigned grid, size_t smem, cudaStream_t stream) {
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(conv3x3_halo_kernel<ACT>, smem));
    launch_pdl(conv3x3_halo_kernel<ACT>, grid, CH_THREADS, smem, stream, ta, tw, ep, g, bn, nt, tiles, na, nw, cols);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

// returns DN_OK after launching, or 1 when the shape is left to the general kernel
static int conv3x3_halo(const void* x, const void* w, const PwEpilogue& ep, int B, int H, int W, int C, int N, cudaStream_t stream) {
    if (W < 8 || H < 2) return 1;
    HaloGeom g{};
    g.B = B, g.H = H, g.W = W, g.C = C, g.N = N;
    // patch: TH * (TW + 2) <= 128 rows; best useful fraction of the 128 GEMM rows, ties -> the smaller window
    double best = -1.0;
    for (int tw = std::min(W, 126); tw >= 6; --tw) {
        const int pw = tw + 2, th = std::min(H, 128 / pw);
        if (th < 1) continue;
        const long long tiles = (long long)ceil_div(W, tw) * ceil_div(H, th);
        const double util = (double)H * W / ((double)tiles * 128.0) - 1e-3 * (double)((th + 2) * pw) / 128.0;
        if (util > best) best = util, g.TH = th, g.TW = tw, g.PW = pw;
    }
    if (best < 0.5) return 1;
    g.tiles_x = ceil_div(W, g.TW), g.tiles_y = ceil_div(H, g.TH);
    g.a_stage_bytes = ((128 + 2 * g.PW + 2) * 128 + 1023) & ~1023;
    const int kblocks = C / C3_BLOCK_K;
    const int nt = ceil_div(N, 256);
    int bn = ceil_div(N, nt);
    bn = nt > 1 ? (bn + 63) & ~63 : (bn + 15) & ~15;
    int cols = 32;
    while (cols < 2 * bn) cols <<= 1;
    const size_t w_tile = (size_t)bn * C3_BLOCK_K * 2;
    const size_t fixed = 1024 + sizeof(HaloBarriers) + ((size_t)nt * bn + 32) * 4;
    const size_t cap = (size_t)(227 * 1024 - C3_STATIC_SMEM - 256);
    int na = 0, nw = 0;
    if (nt == 1 && fixed + 9 * kblocks * w_tile + 2 * (size_t)g.a_stage_bytes <= cap) {
        g.w_resident = 1;
        na = (int)std::min<size_t>(CH_MAX_A, (cap - fixed - 9 * kblocks * w_tile) / g.a_stage_bytes);
    } else {
        na = 2;
        if (fixed + 2 * (size_t)g.a_stage_bytes + 3 * w_tile > cap) return 1;
        nw = (int)std::min<size_t>(CH_MAX_W, (cap - fixed - 2 * (size_t)g.a_stage_bytes) / w_tile);
        while (na < 3 && fixed + (size_t)(na + 1) * g.a_stage_bytes + (size_t)nw * w_tile <= cap) ++na;
    }
    const size_t smem = fixed + (size_t)na * g.a_stage_bytes + (g.w_resident ? 9 * kblocks : nw) * w_tile;
    PFN_encodeTiled3 fn = c3_encode_fn();
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    DN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, DN_ERR_INVALID,
               "dense 3x3: operands must be 16-byte aligned");
    CUtensorMap ta, tw;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)C3_BLOCK_K, (cuuint32_t)g.PW, (cuuint32_t)(g.TH + 2), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = fn(&ta, DN_TMAP_HALF, 4, const_cast<void*>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled (dense 3x3 window) failed (%d)", (int)r);
    }
    int rc = make_tmap_h16_2d(&tw, w, 9ll * N, C, bn, C3_BLOCK_K);
    if (rc) return rc;
    const long long tiles = (long long)g.tiles_x * g.tiles_y * B * nt;
    DN_REQUIRE(tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "dense 3x3 problem too large");
    const unsigned grid = (unsigned)std::min<long long>(sm_count(), tiles);
    switch (ep.act) {
        case DN_ACT_RELU: return ch_launch<DN_ACT_RELU>(ta, tw, ep, g, bn, nt, (int)tiles, na, nw, cols, grid, smem, stream);
        case DN_ACT_RELU6: return ch_launch<DN_ACT_RELU6>(ta, tw, ep, g, bn, nt, (int)tiles, na, nw, cols, grid, smem, stream);
        case DN_ACT_NONE: return ch_launch<DN_ACT_NONE>(ta, tw, ep, g, bn, nt, (int)tiles, na, nw, cols, grid, smem, stream);
    }
    DN_REQUIRE(false, DN_ERR_UNSUPPORTED, "dense 3x3: activation %d not built (none / relu / relu6)", ep.act);
}

int conv3x3_tc(const void* x, const void* w, const PwEpilogue& ep, int B, int H, int W, int C, int N, int stride, int pad, int dil,
               cudaStream_t stream) {
    DN_REQUIRE(C % C3_BLOCK_K == 0, DN_ERR_UNSUPPORTED, "dense 3x3: input channels must be a multiple of 64 (got %d)", C);
    DN_REQUIRE((stride == 1 || stride == 2) && (pad == 0 || pad == dil) && dil >= 1, DN_ERR_UNSUPPORTED,
               "dense 3x3: stride 1 / 2, padding 0 or = dilation");
    ConvGeom g{};
    g.B = B, g.H = H, g.W = W, g.C = C, g.N = N, g.stride = stride, g.pad = pad, g.dil = dil;
    g.Ho = (H + 2 * pad - 2 * dil - 1) / stride + 1;
    g.Wo = (W + 2 * pad - 2 * dil - 1) / stride + 1;
    DN_REQUIRE(g.Ho > 0 && g.Wo > 0, DN_ERR_INVALID, "dense 3x3: empty output");
    if (stride == 1 && dil == 1 && pad == 1 && c3_halo_mode() != 0) {
        const int rc = conv3x3_halo(x, w, ep, B, H, W, C, N, stream);
        if (rc <= 0) return rc;
    }
    c3_pick_patch(B, g.Ho, g.Wo, &g.TB, &g.TH, &g.TW);
    g.tiles_x = ceil_div(g.Wo, g.TW), g.tiles_y = ceil_div(g.Ho, g.TH), g.tiles_b = ceil_div(B, g.TB);
    // N tiling: up to 256 columns per tile (two TMEM buffers of BLOCK_N columns)
    const int nt = ceil_div(N, 256);
    int bn = ceil_div(N, nt);
    bn = nt > 1 ? (bn + 63) & ~63 : (bn + 15) & ~15;
    int cols = 32;
    while (cols < 2 * bn) cols <<= 1;
    auto need = [&](int stg) {
        return 1024 + (size_t)stg * (C3_A_STAGE_BYTES + (size_t)bn * C3_BLOCK_K * 2) + sizeof(C3Barriers) + ((size_t)nt * bn + 32) * 4;
    };
    const size_t cap = (size_t)(227 * 1024 - C3_STATIC_SMEM - 256);
    int st = C3_MAX_STAGES;
    while (st > 2 && need(st) > cap) --st;
    DN_REQUIRE(need(st) <= cap, DN_ERR_UNSUPPORTED, "dense 3x3: tile does not fit in shared memory");
    // tensor maps: A = 4-D NHWC input, box [TB][TH][TW][64] with the convolution's stride as traversal stride; W = [9 N, C]
    PFN_encodeTiled3 fn = c3_encode_fn();
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    DN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, DN_ERR_INVALID,
               "dense 3x3: operands must be 16-byte aligned");
    CUtensorMap ta, tw;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)C3_BLOCK_K, (cuuint32_t)((g.TW - 1) * stride + 1), (cuuint32_t)((g.TH - 1) * stride + 1),
                             (cuuint32_t)g.TB};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = fn(&ta, DN_TMAP_HALF, 4, const_cast<void*>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled (dense 3x3 input) failed (%d)", (int)r);
    }
    int rc = make_tmap_h16_2d(&tw, w, 9ll * N, C, bn, C3_BLOCK_K);
    if (rc) return rc;
    const long long tiles = (long long)g.tiles_x * g.tiles_y * g.tiles_b * nt;
    DN_REQUIRE(tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "dense 3x3 problem too large");
    long long grid = sm_count();
    if (grid > tiles) grid = tiles;
    const size_t smem = need(st);
    switch (ep.act) {
        case DN_ACT_RELU: return c3_launch<DN_ACT_RELU>(ta, tw, ep, g, bn, nt, (int)tiles, st, cols, (unsigned)grid, smem, stream);
        case DN_ACT_RELU6: return c3_launch<DN_ACT_RELU6>(ta, tw, ep, g, bn, nt, (int)tiles, st, cols, (unsigned)grid, smem, stream);
        case DN_ACT_NONE: return c3_launch<DN_ACT_NONE>(ta, tw, ep, g, bn, nt, (int)tiles, st, cols, (unsigned)grid, smem, stream);
    }
    DN_REQUIRE(false, DN_ERR_UNSUPPORTED, "dense 3x3: activation %d not built (none / relu / relu6)", ep.act);
}

}  // namespace dn

using namespace dn;

extern "C" int dn_conv3x3(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C, int N, int stride,
                          int pad, int dilation, int act, int out_fp32, int64_t out_batch_stride, int64_t out_row_stride,
                          void* stream_) {
    DN_REQUIRE(x && w && bias && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && N > 0, DN_ERR_INVALID, "bad shape");
    const int Ho = (H + 2 * pad - 2 * dilation - 1) / stride + 1, Wo = (W + 2 * pad - 2 * dilation - 1) / stride + 1;
    PwEpilogue ep;
    ep.bias = bias;
    ep.residual = nullptr;
    ep.y = y;
    ep.N = N;
    ep.act = act;
    ep.out_fp32 = out_fp32;
    ep.hw = Ho * Wo;
    ep.out_batch_stride = out_batch_stride ? out_batch_stride : (long long)Ho * Wo * N;
    ep.out_row_stride = out_row_stride ? out_row_stride : N;
    return conv3x3_tc(x, w, ep, B, H, W, C, N, stride, pad, dilation, (cudaStream_t)stream_);
}
