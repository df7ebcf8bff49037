// Pointwise (1x1) convolution as a tcgen05 tensor-core GEMM for sm_100a.
//
//   D[M, N] = A[M, K] . W[N, K]^T     A = NHWC bf16 activations (K-major), W = bf16 weights (K-major)
//
// Persistent CTAs (one to four per SM) walk 128 x BLOCK_N output tiles (BLOCK_N is a runtime multiple
// of 16, <= 256); the smem ring and two TMEM accumulator buffers run continuously across tiles:
//   warp 0      TMA producer : cp.async.bulk.tensor 2D loads of the A and W k-blocks (64 bf16 = one
//                              128-byte swizzle row) into a NUM_STAGES-deep shared-memory ring,
//                              completion signalled on mbarriers (expect_tx)
//   warp 1      MMA issuer   : allocates TMEM, one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                              (UMMA 128 x BLOCK_N x 16, fp32 accumulator in TMEM), tcgen05.commit frees
//                              the smem stage / publishes the accumulator
//   warps 2..5  epilogue     : tcgen05.ld the accumulator (thread = output row), add the folded-BN bias and
//                              apply the activation, stage 32-column fp32 chunks through shared memory so
//                              that the residual loads / output stores are row-contiguous (8 lanes per row),
//                              add the residual, convert, store with the head's strided addressing
// K and N tails are zero-filled by TMA (out-of-bounds box elements), M tails are masked in the epilogue.
// Reference ops replaced: see dn_pwconv in include/demonet_b200.h.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "pwconv.cuh"
#include "tcgen05.cuh"

namespace dn {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;                 // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 16;
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_RING_MAX = 8;                  // barrier slots; the cta_group::2 form (half-size weight stages) rings deeper
constexpr int TC_EPI_WARPS = 8;                 // two warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_SCALE_WARPS = TC_EPI_WARPS;    // SCALE_A variant: the epilogue warps also rescale the A tiles in shared memory
constexpr int TC_A_STAGE_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;      // 16 KiB
constexpr int TC_STAGE_PITCH = 36;                                 // floats; 144-B rows keep 16-B smem accesses conflict-free
constexpr int TC_STAGE_BYTES = 5120;                               // >= 32*36*4 and a multiple of 1024
constexpr int TC_STATIC_SMEM = TC_EPI_WARPS * TC_STAGE_BYTES + 1024;      // epilogue staging (+ alignment)

struct __align__(8) TcBarriers {
    uint64_t full[TC_RING_MAX];        // TMA -> MMA: k-block landed in smem
    uint64_t empty[TC_RING_MAX];       // MMA -> TMA: smem stage consumed
    uint64_t tmem_full[2];               // MMA -> epilogue: accumulator buffer complete
    uint64_t tmem_empty[2];              // epilogue -> MMA: accumulator buffer drained
    uint64_t scaled[TC_RING_MAX];      // SCALE_A: scaler warps -> MMA: the A tile of the stage has been rescaled
    uint64_t w_full;                     // W-stationary mode: the CTA's weight tile (all k-blocks) has landed
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent kernel: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the smem ring and
// the two TMEM accumulator buffers run continuously across tiles, so the TMA loads / MMAs of tile i+1
// overlap the epilogue of tile i.
template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));      // x * relu6(x + 3) / 6
    return v;
}

// SCALE_A (squeeze-excitation folded into the project GEMM): before they wait for the accumulator of a tile, the eight
// epilogue warps (idle during the K loop anyway) multiply the A tile of every stage, in shared memory, by the
// per-(image, input channel) scale -- a[m][k] <- round16(a[m][k] * scale[m / hw][k]), exactly what se_scale_kernel writes
// to HBM in the unfused path -- between the TMA completion and the MMA issue.  The scaled tensor never exists in HBM:
// one read and one write of every SE tensor disappear.  (The K loop of tile i + 1 then starts behind the epilogue of
// tile i; these GEMMs have long K and narrow N, and the other CTAs of the SM fill the gap.)
// CG2: the cta_group::2 form of PAIR mode (pair == 2).  A kernel that contains cta_group::2 instructions can only be launched
// as clusters ("cluster misconfiguration" otherwise), so that code lives in its own instantiations.
template <int ACT, bool TMA_STORE, bool SCALE_A, bool CG2 = false>
__global__ void __launch_bounds__(TC_THREADS, 2)
pwconv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_y, PwEpilogue ep,
                 int M, int K, int N, int block_n, int n_tiles, int num_tiles, int num_stages, int tmem_cols, int w_stat, int pair) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // per epilogue warp: 32 rows x 32 fp32 (+pad) for the staged stores, or a 32 x 64 bf16 SWIZZLE_128B tile
    // for the TMA store
    __shared__ __align__(1024) uint8_t s_stage_raw[TC_EPI_WARPS][TC_STAGE_BYTES];
    // carve: [stages x A tile][stages x W tile][barriers]; tiles must be 1024-B aligned for SWIZZLE_128B
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space
    const int w_stage_bytes = (CG2 ? block_n >> 1 : block_n) * TC_BLOCK_K * 2;      // cta_group::2: a CTA only holds its half of B
    uint8_t* smem_a = smem;
    // W-stationary mode (w_stat): the grid is a multiple of n_tiles, so a CTA meets ONE weight tile for its whole life; all
    // its k-blocks are loaded once into their own buffers and the ring only carries A.  Otherwise the (L2-resident)
    // weight tile is fetched again with every output tile -- for the expand layers (K <= 128, N up to 256 per tile) that
    // is more bytes into the SM than the activations themselves.
    const int num_k_blocks = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;
    uint8_t* smem_w = smem + num_stages * TC_A_STAGE_BYTES;
    TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem_w + (w_stat ? num_k_blocks : num_stages) * w_stage_bytes);
    // the whole (zero-padded) bias vector, staged once per CTA: weights are not produced by the previous kernel, so this
    // happens before the programmatic-dependency wait, and no epilogue chunk waits for a global load any more
    float* s_bias_all = reinterpret_cast<float*>(bars + 1);            // [n_tiles * block_n + 32]
    for (int i = threadIdx.x; i < n_tiles * block_n + 32; i += blockDim.x) s_bias_all[i] = (i < N) ? __ldg(ep.bias + i) : 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_w);
        if (TMA_STORE) prefetch_tmap(&tmap_y);
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], pair == 1 ? 2 : 1);              // multicast pair mode: both CTAs' MMAs release a stage
            mbar_init(&bars->scaled[s], TC_SCALE_WARPS);            // one arrival per scaler warp (SCALE_A only)
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->tmem_full[b], 1);
            mbar_init(&bars->tmem_empty[b], CG2 ? 2 * TC_EPI_WARPS : TC_EPI_WARPS);      // one arrival per epilogue warp (of both CTAs)
        }
        mbar_init(&bars->w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (CG2) tmem_alloc_pair(&bars->tmem_base, (uint32_t)tmem_cols);
        else tmem_alloc(&bars->tmem_base, (uint32_t)tmem_cols);
    }
    pdl_trigger();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    // PAIR mode (launched as clusters of two CTAs; K > 128 layers, whose weight tile is re-fetched with every output tile
    // and, with the A tile, makes the L2 -> shared-memory operand stream the bound of the kernel): the two CTAs of a
    // cluster work on two M tiles of the SAME N tile in lockstep, each loads HALF of every weight k-block and multicasts
    // it into both CTAs' stage (one L2 read serves two SMs), and a stage is released by both CTAs' MMAs (multicast commit
    // on both `empty` barriers).  `num_tiles` then counts pair tiles: (pairs of M tiles) x N tiles.
    const uint32_t rank = pair ? cluster_ctarank() : 0u;
    if (pair) cluster_sync_all();                     // the peer's barriers are initialised before anything lands on them
    const uint32_t tmem_base = bars->tmem_base;       // (read behind the cluster barrier: a cta_group::2 allocation is collective)
    const int t_first = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int t_step = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto tile_m0 = [&](int tile) {
        return (pair ? (tile / n_tiles) * 2 + (int)rank : tile / n_tiles) * TC_BLOCK_M;
    };
    pdl_wait();                 // descriptors prefetched, barriers and TMEM set up while the previous kernel drains

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            const uint32_t stage_bytes = (uint32_t)(TC_A_STAGE_BYTES + (w_stat ? 0 : w_stage_bytes));
            if (w_stat) {
                const int n0 = ((int)blockIdx.x % n_tiles) * block_n;
                mbar_expect_tx(&bars->w_full, (uint32_t)(num_k_blocks * w_stage_bytes));
                for (int kb = 0; kb < num_k_blocks; ++kb)
                    tma_load_2d(smem_w + kb * w_stage_bytes, &tmap_w, &bars->w_full, kb * TC_BLOCK_K, n0);
            }
            uint32_t it = 0;
            for (int tile = t_first; tile < num_tiles; tile += t_step) {
                const int m0 = tile_m0(tile), n0 = (tile % n_tiles) * block_n;
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int s = it % num_stages;
                    mbar_wait_producer(&bars->empty[s], ((it / num_stages) & 1u) ^ 1u);
                    if constexpr (CG2) {
                        // cta_group::2: my A rows and my HALF of the weight k-block stay in my shared memory; the bytes of
                        // both CTAs complete on the leader's barrier, which the leader arms for both
                        if (rank == 0) mbar_expect_tx(&bars->full[s], 2u * (uint32_t)(TC_A_STAGE_BYTES + w_stage_bytes));
                        tma_load_2d_pair(smem_a + s * TC_A_STAGE_BYTES, &tmap_a, &bars->full[s], kb * TC_BLOCK_K, m0);
                        tma_load_2d_pair(smem_w + s * w_stage_bytes, &tmap_w, &bars->full[s], kb * TC_BLOCK_K,
                                         n0 + (int)rank * (block_n >> 1));
                        continue;
                    }
                    mbar_expect_tx(&bars->full[s], stage_bytes);
                    tma_load_2d(smem_a + s * TC_A_STAGE_BYTES, &tmap_a, &bars->full[s], kb * TC_BLOCK_K, m0);
                    if (pair)           // my half of the weight k-block (tmap_w's box is block_n / 2 rows) into both CTAs
                        tma_load_2d_multicast(smem_w + s * w_stage_bytes + rank * (uint32_t)(w_stage_bytes >> 1), &tmap_w, &bars->full[s],
                                              kb * TC_BLOCK_K, n0 + (int)rank * (block_n >> 1), (uint16_t)3);
                    else if (!w_stat) tma_load_2d(smem_w + s * w_stage_bytes, &tmap_w, &bars->full[s], kb * TC_BLOCK_K, n0);
                }
            }
        }
    } else if (warp == 1 && !(CG2 && rank != 0)) {
        // ===== MMA issuer (cta_group::2: the leader CTA issues 256 x block_n UMMAs for the pair) =====
        const uint32_t idesc = make_idesc(CG2 ? 2 * TC_BLOCK_M : TC_BLOCK_M, block_n);
        uint32_t it = 0, lt = 0;
        if (w_stat) mbar_wait(&bars->w_full, 0u);
        for (int tile = t_first; tile < num_tiles; tile += t_step, ++lt) {
            const uint32_t buf = lt & 1u;
            mbar_wait(&bars->tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);      // epilogue has drained this buffer
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * (uint32_t)block_n;
            for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                const int s = it % num_stages;
                mbar_wait(SCALE_A ? &bars->scaled[s] : &bars->full[s], (it / num_stages) & 1u);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_smem_desc(smem_u32(smem_a + s * TC_A_STAGE_BYTES));
                    const uint64_t dw = make_smem_desc(smem_u32(smem_w + (w_stat ? kb : s) * w_stage_bytes));
                    const int k_left = K - kb * TC_BLOCK_K;              // zero-filled K tail needs no MMA
                    const int ksteps = k_left >= TC_BLOCK_K ? TC_BLOCK_K / TC_UMMA_K : (k_left + TC_UMMA_K - 1) / TC_UMMA_K;
                    for (int k = 0; k < ksteps; ++k) {
                        // advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in (addr >> 4) units
                        if constexpr (CG2) umma_f16_pair(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_f16(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // frees this smem stage when the MMAs retire (pair modes: in both CTAs)
                    if constexpr (CG2) umma_commit_pair(&bars->empty[s], (uint16_t)3);
                    else if (pair) umma_commit_multicast(&bars->empty[s], (uint16_t)3);
                    else umma_commit(&bars->empty[s]);
                    if (kb == num_k_blocks - 1) {                              // accumulator complete (cta_group::2: in both CTAs)
                        if constexpr (CG2) umma_commit_pair(&bars->tmem_full[buf], (uint16_t)3);
                        else umma_commit(&bars->tmem_full[buf]);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= 2) {
        // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column chunks interleaved between the
        // two warps that share a quarter =====
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        uint32_t lt = 0, stores = 0, sit = 0;
        // SCALE_A: rescale the A tiles of output tile `tile_s` (see the kernel comment).  Software-pipelined one tile ahead
        // of the epilogue: the warps rescale tile i + 1 BEFORE they wait for the accumulator of tile i, so the MMAs of
        // tile i + 1 (and the global loads of its scales) run under the epilogue of tile i instead of in front of it.
        auto scale_tile = [&](int tile_s) {
            const int m0 = tile_m0(tile_s);
                // ---- A-operand scaling.  A stage holds 128 rows of 128 bytes (SWIZZLE_128B: the 16-byte chunk at position
                // p of row r carries the k-chunk p ^ (r & 7)).  Warp w owns rows 16 w .. 16 w + 15; one warp instruction
                // covers 4 rows x 8 chunks = 512 contiguous bytes (conflict-free), four instructions per stage.
                constexpr int SC_ROWS = TC_BLOCK_M / TC_SCALE_WARPS, SC_IT = SC_ROWS / 4;
                const int sw = warp - 2, p = lane & 7, rsub = lane >> 3;
                int bimg[SC_IT];
#pragma unroll
                for (int i = 0; i < SC_IT; ++i) {
                    const int mr = m0 + sw * SC_ROWS + i * 4 + rsub;
                    bimg[i] = (mr < M) ? mr / ep.hw : -1;             // rows past M are zero-filled by TMA: nothing to scale
                }
                // rows of one image share their scales, and a lane only ever meets two k-chunks per k-block (p ^ rsub and
                // p ^ (rsub + 4)): when its rows lie inside one image the scales are loaded once per k-block, ahead of the
                // barrier wait (rows grow with i, so first == last means every row; the vote keeps the branch uniform)
                const bool one_image = __all_sync(0xffffffffu, bimg[0] >= 0 && bimg[0] == bimg[SC_IT - 1]);
                for (int kb = 0; kb < num_k_blocks; ++kb, ++sit) {
                    const int s = sit % num_stages;
                    float4 sc[2][2];
                    if (one_image) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int k = kb * TC_BLOCK_K + ((p ^ (rsub + 4 * h)) << 3);
                            const float* sp = ep.a_scale + (long long)bimg[0] * ep.a_scale_c + (k % ep.a_scale_c);
                            sc[h][0] = (k < K) ? __ldg(reinterpret_cast<const float4*>(sp)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            sc[h][1] = (k < K) ? __ldg(reinterpret_cast<const float4*>(sp) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    mbar_wait(&bars->full[s], (sit / num_stages) & 1u);
                    uint8_t* a_tile = smem_a + s * TC_A_STAGE_BYTES;
                    if (one_image) {
#pragma unroll
                        for (int i = 0; i < SC_IT; ++i) {
                            const int r = sw * SC_ROWS + i * 4 + rsub;      // r & 7 == rsub + 4 * (i & 1)
                            uint4* q = reinterpret_cast<uint4*>(a_tile + r * 128 + (p << 4));
                            const float4 s0 = sc[i & 1][0], s1 = sc[i & 1][1];
                            float f[8];
                            unpack8(*q, f);
                            f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
                            f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
                            *q = pack8(f);                                  // zero-filled K tail: 0 * 0 = 0
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < SC_IT; ++i) {
                            const int r = sw * SC_ROWS + i * 4 + rsub;
                            const int k = kb * TC_BLOCK_K + ((p ^ (r & 7)) << 3);
                            if (bimg[i] >= 0 && k < K) {                // K % 8 == 0: a chunk is inside K or entirely in the zero tail
                                const float* sp = ep.a_scale + (long long)bimg[i] * ep.a_scale_c + (k % ep.a_scale_c);
                                const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp));
                                const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp) + 1);
                                uint4* q = reinterpret_cast<uint4*>(a_tile + r * 128 + (p << 4));
                                float f[8];
                                unpack8(*q, f);
                                f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
                                f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
                                *q = pack8(f);
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->scaled[s]);
                }
        };
        if constexpr (SCALE_A) {
            if (t_first < num_tiles) scale_tile(t_first);
        }
        for (int tile = t_first; tile < num_tiles; tile += t_step, ++lt) {
            const uint32_t buf = lt & 1u;
            const int m0 = tile_m0(tile), n0 = (tile % n_tiles) * block_n;
            const int m = m0 + row;
            if constexpr (SCALE_A) {
                if (tile + t_step < num_tiles) scale_tile(tile + t_step);
            }
            const int n_valid = min(block_n, N - n0);
            mbar_wait(&bars->tmem_full[buf], (lt >> 1) & 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * (uint32_t)block_n + ((uint32_t)(quarter * 32) << 16);
            if constexpr (TMA_STORE) {
                // bf16 output, no residual: pack 32-column groups into SWIZZLE_64B staging tiles (thread = row;
                // 16-byte chunk c of row r lives at chunk position c ^ ((r >> 1) & 3)) and let the TMA engine
                // write them out; M / N tails are clipped by the tensor map.  Two staging tiles per warp.
                uint8_t* obase = s_stage_raw[warp - 2];
                for (int c0 = half * 32; c0 < n_valid; c0 += 64, ++stores) {
                    uint8_t* obuf = obase + (stores & 1u) * 2048;
                    const float* sbw = s_bias_all + n0 + c0;
                    uint32_t v[32];
                    const bool second = (c0 + 16 < block_n);              // warp-uniform
                    tmem_ld16(tmem_d + (uint32_t)c0, v);
                    if (second) tmem_ld16(tmem_d + (uint32_t)(c0 + 16), v + 16);
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // this tile's previous store
                    __syncwarp();
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o;
                        if constexpr (ACT == DN_ACT_HSWISH) {
                            float f[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = act_fn<ACT>(__uint_as_float(v[q * 8 + i]) + sbw[q * 8 + i]);
                            o = pack8(f);
                        } else {
                            // bias add as packed FFMA2 (x * 1 + b: one rounding, = the fp32 add), ReLU / ReLU6 on the packed
                            // bf16 pair AFTER the rounding (rounding is monotone and keeps 0 and 6: same values)
                            uint32_t w[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 bb = *reinterpret_cast<const float2*>(sbw + q * 8 + 2 * i);
                                const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 2 * i]), __uint_as_float(v[q * 8 + 2 * i + 1])),
                                                            make_float2(1.f, 1.f), bb);
                                dn_half2_t h = floats_to_half2(t.x, t.y);
                                if constexpr (ACT == DN_ACT_RELU || ACT == DN_ACT_RELU6) h = __hmax2(h, half2_const(0.f));
                                if constexpr (ACT == DN_ACT_RELU6) h = __hmin2(h, half2_const(6.f));
                                w[i] = *reinterpret_cast<uint32_t*>(&h);
                            }
                            o = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                        *reinterpret_cast<uint4*>(obuf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = o;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_y, obuf, n0 + c0, m0 + quarter * 32);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else {
            const long long out_row = (m < M) ? ep.row_offset(m) : 0;     // this lane's own row; shared by shuffle below
            float* stg = reinterpret_cast<float*>(s_stage_raw[warp - 2]);
            for (int c0 = half * 32; c0 < n_valid; c0 += 64) {
                uint32_t v[32];
                const bool second = (c0 + 16 < block_n);                  // warp-uniform
                tmem_ld16(tmem_d + (uint32_t)c0, v);
                if (second) tmem_ld16(tmem_d + (uint32_t)(c0 + 16), v + 16);
                const float* sbw = s_bias_all + n0 + c0;
                tmem_ld_wait();
                // residual tile of this chunk: issue all loads now, they complete behind the TMEM read / staging
                const int cg = (lane & 7) * 4;
                const int n = n0 + c0 + cg;
                const int cnt = n0 + n_valid - n;                          // columns of THIS tile left in the row (>= 4: full vector)
                const bool res_vec = ep.residual && cnt >= 4 && (N % 4 == 0);
                uint2 rres[8];
                if (res_vec) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int mr = m0 + quarter * 32 + it * 4 + (lane >> 3);
                        rres[it] = (mr < M) ? __ldg(reinterpret_cast<const uint2*>(ep.residual + (long long)mr * N + n))
                                            : make_uint2(0u, 0u);
                    }
                }
                // (1) thread = row: bias + activation in fp32, staged to shared memory (pitch 36 floats)
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = act_fn<ACT>(__uint_as_float(v[j + 0]) + sbw[j + 0]);
                    o.y = act_fn<ACT>(__uint_as_float(v[j + 1]) + sbw[j + 1]);
                    o.z = act_fn<ACT>(__uint_as_float(v[j + 2]) + sbw[j + 2]);
                    o.w = act_fn<ACT>(__uint_as_float(v[j + 3]) + sbw[j + 3]);
                    *reinterpret_cast<float4*>(stg + lane * TC_STAGE_PITCH + j) = o;
                }
                __syncwarp();
                // (2) 8 lanes per row, 4 rows per instruction: coalesced output stores
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + (lane >> 3);
                    const long long orow = __shfl_sync(0xffffffffu, out_row, r);
                    const int mr = m0 + quarter * 32 + r;
                    if (mr < M && cnt > 0) {
                        float4 o = *reinterpret_cast<const float4*>(stg + r * TC_STAGE_PITCH + cg);
                        float f[4] = {o.x, o.y, o.z, o.w};
                        if (res_vec) {
                            const float2 r0 = h2_to_float2(rres[it].x), r1 = h2_to_float2(rres[it].y);
                            f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
                        } else if (ep.residual) {
                            const dn_half_t* rp = ep.residual + (long long)mr * N + n;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (i < cnt) f[i] += half_to_float(rp[i]);
                        }
                        if (ep.out_fp32) {
                            float* dst = reinterpret_cast<float*>(ep.y) + orow + n;
                            if (cnt >= 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                                *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
                            } else if (cnt >= 4 && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
                                reinterpret_cast<float2*>(dst)[0] = make_float2(f[0], f[1]);
                                reinterpret_cast<float2*>(dst)[1] = make_float2(f[2], f[3]);
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
                __syncwarp();
            }
            }
            // all tcgen05.ld of this warp have completed (wait::ld above): hand the buffer back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG2) mbar_arrive_leader(&bars->tmem_empty[buf]);
                else mbar_arrive(&bars->tmem_empty[buf]);
            }
        }
        if (TMA_STORE && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();                     // neither CTA leaves while the other may still signal its barriers
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (CG2) tmem_dealloc_pair(tmem_base, (uint32_t)tmem_cols);
        else tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
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

// 2D bf16 tensor map over a row-major [rows, cols] matrix, box = [box_rows, box_cols]; the swizzle span equals
// the box row (64 cols -> SWIZZLE_128B, 32 -> SWIZZLE_64B, 16 -> SWIZZLE_32B)
int make_tmap_h16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, int box_cols) {
    PFN_encodeTiled fn = get_encode_fn();
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    DN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, DN_ERR_INVALID, "GEMM operand must be 16-byte aligned");
    DN_REQUIRE(cols % 8 == 0, DN_ERR_UNSUPPORTED, "GEMM K must be a multiple of 8 (got %lld)", cols);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    DN_REQUIRE(box_cols == 16 || box_cols == 32 || box_cols == 64, DN_ERR_INVALID, "box_cols must be 16, 32 or 64");
    const CUtensorMapSwizzle swz = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, DN_TMAP_HALF, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d", (int)r,
               rows, cols, box_rows);
    return DN_OK;
}

// Largest dynamic shared-memory request that still lets n CTAs co-reside on an SM (228 KiB per SM;
// every CTA costs its dynamic request + the static staging buffers + 1 KiB reserved).
static size_t smem_cap(int n) {
    return n == 1 ? (size_t)(227 * 1024 - TC_STATIC_SMEM - 256) : (size_t)(228 * 1024) / n - TC_STATIC_SMEM - 1024 - 256;
}

void pwconv_tc_plan(long long m_plan, int K, int N, int* block_n, int* n_tiles, int* stages, int* tmem_cols, size_t* smem_bytes,
                    int* w_stationary, int* pair, int pair_force) {
    static const int bn_max = [] {                     // measurement aid: DN_PW_BN_MAX=64|128 caps the tile width
        const char* v = getenv("DN_PW_BN_MAX");
        const int x = v ? atoi(v) : 256;
        return (x == 64 || x == 128) ? x : 256;
    }();
    // small maps: with 256-wide tiles the grid would leave most SMs idle, so split N further (r01 A/B)
    const long long m_tiles = (m_plan + TC_BLOCK_M - 1) / TC_BLOCK_M;
    int cap_n = bn_max;
    while (cap_n > 64 && m_tiles * ((N + cap_n - 1) / cap_n) < sm_count()) cap_n >>= 1;
    const int nt = (N + cap_n - 1) / cap_n;
    int bn = (N + nt - 1) / nt;
    bn = nt > 1 ? (bn + 63) & ~63 : (bn + 15) & ~15;      // 64-column groups must not straddle two N tiles
    int cols = 32;
    while (cols < 2 * bn) cols <<= 1;                  // two accumulator buffers
    *block_n = bn;
    *n_tiles = nt;
    const int kb = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;
    int st = kb >= 3 ? TC_MAX_STAGES : kb + 1;            // short-K layers: fewer stages -> more CTAs per SM
    int st_max = kb + 1;
    // W-stationary: the weight tile of a CTA (all k-blocks) stays in shared memory and the ring carries A only
    // (DN_PW_WSTAT=0 turns it off, measurement aid)
    static const int wstat_on = [] {
        const char* v = getenv("DN_PW_WSTAT");
        return v ? atoi(v) : 1;
    }();
    // (only for K <= 128, i.e. at most two k-blocks: the resident tile then never takes more shared memory than the
    // weight share of the two-stage ring it replaces, so the CTA count per SM cannot drop)
    const size_t w_total = (size_t)kb * bn * TC_BLOCK_K * 2;
    const bool ws = wstat_on && kb <= 2;
    if (w_stationary) *w_stationary = ws ? 1 : 0;
    if (ws) st = st_max = TC_MAX_STAGES;                   // A-only stages: the ring can run several tiles ahead
    // PAIR modes (DN_PW_PAIR=1|2, read per call; OFF by default), for layers whose weight tile travels through the ring
    // (K > 128) and is wide enough to matter, on maps with at least two M tiles per SM:
    //   1  clusters of two CTAs, every weight k-block loaded half by each CTA and TMA-multicast into both ring stages
    //   2  cta_group::2: the pair's leader issues 256 x block_n UMMAs, each CTA keeps only ITS half of B in shared memory
    //      (half-size weight stages -> a deeper ring in the same shared memory)
    const char* pv = getenv("DN_PW_PAIR");
    const int pair_env = pair_force >= 0 ? pair_force : (pv ? atoi(pv) : 0);      // pair_force: the mode a caller already committed to
    const int pm = ((pair_env == 1 || pair_env == 2) && !ws && bn >= 64 && m_tiles >= 2 * (long long)sm_count()) ? pair_env : 0;
    if (pair) *pair = pm;
    const size_t w_stage = (size_t)(pm == 2 ? bn / 2 : bn) * TC_BLOCK_K * 2;
    if (pm == 2) st = st_max = TC_RING_MAX;
    auto need = [&](int stg) {
        const size_t ring = ws ? (size_t)stg * TC_A_STAGE_BYTES + w_total : (size_t)stg * (TC_A_STAGE_BYTES + w_stage);
        return 1024 + ring + sizeof(TcBarriers) + ((size_t)nt * bn + 32) * 4;
    };
    while (st > 2 && need(st) > smem_cap(1)) --st;
    // Resident CTAs beat ring depth: these GEMMs are bound by the latency chains of the epilogue warps and of the
    // load -> MMA -> commit loop, not by bytes in flight (same total either way), so take the largest CTA count whose
    // TMEM columns and a 2-stage ring fit, then deepen the ring as far as that count allows (r01 A/B: never slower,
    // up to 1.45x on the K >> N project layers).  DN_PW_OCC=0 restores the deep-ring plan (measurement aid).
    static const int prefer_occ = [] {
        const char* v = getenv("DN_PW_OCC");
        return v ? atoi(v) : 1;
    }();
    if (prefer_occ) {
        for (int n = 4; n >= 2; --n)
            if (n * cols <= 512 && need(2) <= smem_cap(n)) {
                int s2 = 2;
                while (s2 < (pm == 2 ? TC_RING_MAX : TC_MAX_STAGES) && s2 < st_max && need(s2 + 1) <= smem_cap(n)) ++s2;
                st = s2;
                break;
            }
    }
    *stages = st;
    *tmem_cols = cols;
    *smem_bytes = need(st);
}

template <int ACT, bool TMA_STORE, bool SCALE_A = false>
static int launch_variant(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& ty, const PwEpilogue& ep, int M,
                          int K, int N, int bn, int nt, int tiles, int st, int cols, unsigned grid, size_t smem_req,
                          cudaStream_t stream, int w_stat = 0, int pair = 0) {
    if constexpr (!SCALE_A) {
        if (pair == 2) {
            static SmemOptIn optin2;
            DN_CHECK_CUDA(optin2.ensure(pwconv_tc_kernel<ACT, TMA_STORE, false, true>, smem_cap(1)));
            launch_pdl_cluster(pwconv_tc_kernel<ACT, TMA_STORE, false, true>, grid, TC_THREADS, smem_req, stream, 2u, ta, tw, ty, ep, M,
                               K, N, bn, nt, tiles, st, cols, w_stat, pair);
            DN_CHECK_LAUNCH();
            return DN_OK;
        }
    }
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(pwconv_tc_kernel<ACT, TMA_STORE, SCALE_A>, smem_cap(1)));
    if (pair)
        launch_pdl_cluster(pwconv_tc_kernel<ACT, TMA_STORE, SCALE_A>, grid, TC_THREADS, smem_req, stream, 2u, ta, tw, ty, ep, M, K, N,
                           bn, nt, tiles, st, cols, w_stat, pair);
    else
        launch_pdl(pwconv_tc_kernel<ACT, TMA_STORE, SCALE_A>, grid, TC_THREADS, smem_req, stream, ta, tw, ty, ep, M, K, N, bn, nt,
                   tiles, st, cols, w_stat, pair);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

// ty: tensor map of the bf16 output ([rows, N], box 32 x 64) or nullptr.  The TMA-store epilogue is used when
// ty is given and the layer has neither a residual nor fp32 / strided output.
int pwconv_tc_launch(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap* ty, const PwEpilogue& ep, int M, long long m_plan, int K,
                     int N, cudaStream_t stream, int pair_planned) {
    int bn, nt, st, cols, ws, pair;
    size_t smem;
    // pair_planned: the mode `tw` was built for (its box is block_n / 2 rows in the pair modes).  A scaling is per CTA, so the
    // SE GEMMs take the multicast form where cta_group::2 was asked for (same weight map, but full-size weight stages: plan again)
    pwconv_tc_plan(m_plan, K, N, &bn, &nt, &st, &cols, &smem, &ws, &pair, pair_planned);
    if (pair == 2 && ep.a_scale) pwconv_tc_plan(m_plan, K, N, &bn, &nt, &st, &cols, &smem, &ws, &pair, 1);
    // Resident CTAs per SM: limited by shared memory and by TMEM columns (512 per SM).  The dynamic request is
    // padded up to the largest size that still lets `per_sm` CTAs co-reside, so that the hardware cannot place
    // one more (a CTA that cannot get its TMEM columns would spin until a neighbour exits).
    DN_REQUIRE(smem <= smem_cap(1), DN_ERR_UNSUPPORTED, "GEMM tile does not fit in shared memory");
    int per_sm = 1;
    for (int n = 4; n >= 1; --n)
        if (smem <= smem_cap(n) && n * cols <= 512) {
            per_sm = n;
            break;
        }
    const size_t smem_req = smem_cap(per_sm);
    long long tiles = (long long)ceil_div(M, TC_BLOCK_M) * nt;
    DN_REQUIRE(tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "GEMM too large");
    long long grid = (long long)sm_count() * per_sm;
    if (pair) {                  // pair tiles: two M tiles of one N tile per cluster step; an even grid of whole clusters
        tiles = (long long)ceil_div(ceil_div(M, TC_BLOCK_M), 2) * nt;
        grid &= ~1ll;
        if (grid > 2 * tiles) grid = 2 * tiles;
    } else if (grid > tiles) grid = tiles;
    if (ws) {                    // a CTA must keep meeting the same weight tile: tile % nt == blockIdx.x % nt
        grid = grid / nt * nt;
        if (grid < nt) ws = 0, grid = std::min<long long>((long long)sm_count() * per_sm, tiles);
    }
    const bool tma_store = ty != nullptr && ep.residual == nullptr && !ep.out_fp32;
    const CUtensorMap& tyr = tma_store ? *ty : ta;
    if (ep.a_scale) {           // squeeze-excitation folded into its project GEMM (no activation there: mobilenetv3.py:88-89)
        DN_REQUIRE(ep.act == DN_ACT_NONE && ep.a_scale_c > 0 && ep.a_scale_c % 8 == 0 && K % ep.a_scale_c == 0, DN_ERR_UNSUPPORTED,
                   "A-operand scaling needs a linear GEMM whose K is a multiple of the scale width");
        return tma_store ? launch_variant<DN_ACT_NONE, true, true>(ta, tw, tyr, ep, M, K, N, bn, nt, (int)tiles, st, cols,
                                                                   (unsigned)grid, smem_req, stream, ws, pair)
                         : launch_variant<DN_ACT_NONE, false, true>(ta, tw, tyr, ep, M, K, N, bn, nt, (int)tiles, st, cols,
                                                                    (unsigned)grid, smem_req, stream, ws, pair);
    }
#define DN_PW_CASE(ACT)                                                                                                   \
    return tma_store ? launch_variant<ACT, true>(ta, tw, tyr, ep, M, K, N, bn, nt, (int)tiles, st, cols, (unsigned)grid,  \
                                                 smem_req, stream, ws, pair)                                              \
                     : launch_variant<ACT, false>(ta, tw, tyr, ep, M, K, N, bn, nt, (int)tiles, st, cols, (unsigned)grid, \
                                                  smem_req, stream, ws, pair)
    switch (ep.act) {
        case DN_ACT_RELU: DN_PW_CASE(DN_ACT_RELU);
        case DN_ACT_RELU6: DN_PW_CASE(DN_ACT_RELU6);
        case DN_ACT_HSWISH: DN_PW_CASE(DN_ACT_HSWISH);
        default: DN_PW_CASE(DN_ACT_NONE);
    }
#undef DN_PW_CASE
}

int pwconv_tc(const void* x, const void* w, const PwEpilogue& ep, int M, int K, int N, cudaStream_t stream) {
    int bn, nt, st, cols, pair;
    size_t smem;
    pwconv_tc_plan(M, K, N, &bn, &nt, &st, &cols, &smem, nullptr, &pair);
    CUtensorMap ta, tw, ty;
    int rc = make_tmap_h16_2d(&ta, x, M, K, TC_BLOCK_M, TC_BLOCK_K);
    if (rc) return rc;
    rc = make_tmap_h16_2d(&tw, w, N, K, pair ? bn / 2 : bn, TC_BLOCK_K);
    if (rc) return rc;
    const bool dense = !ep.out_fp32 && !ep.residual && N % 8 == 0 && ep.out_row_stride == N &&
                       (ep.hw >= M || ep.out_batch_stride == (long long)ep.hw * N);
    if (dense) {
        rc = make_tmap_h16_2d(&ty, ep.y, M, N, 32, 32);
        if (rc) return rc;
    }
    return pwconv_tc_launch(ta, tw, dense ? &ty : nullptr, ep, M, M, K, N, stream, pair);
}

}  // namespace dn

using namespace dn;

extern "C" int dn_pwconv_plan_info(long long M, int K, int N, int32_t* out8) {
    DN_REQUIRE(out8 != nullptr && M > 0 && K > 0 && N > 0 && K % 8 == 0, DN_ERR_INVALID, "bad GEMM shape");
    int bn, nt, st, cols, ws, pair;
    size_t smem;
    pwconv_tc_plan(M, K, N, &bn, &nt, &st, &cols, &smem, &ws, &pair);
    int per_sm = 0;
    for (int n = 4; n >= 1; --n)
        if (smem <= smem_cap(n) && n * cols <= 512) {
            per_sm = n;
            break;
        }
    out8[0] = bn, out8[1] = nt, out8[2] = st, out8[3] = cols, out8[4] = (int32_t)smem, out8[5] = ws, out8[6] = pair, out8[7] = per_sm;
    return DN_OK;
}

extern "C" int dn_pwconv(const void* x, const void* w, const float* bias, const void* residual, void* y, int M, int K,
                         int N, int act, int out_fp32, int hw, int64_t out_batch_stride, int64_t out_row_stride, int impl,
                         void* stream_) {
    DN_REQUIRE(x && w && bias && y, DN_ERR_INVALID, "NULL tensor pointer");
    DN_REQUIRE(M > 0 && K > 0 && N > 0 && hw > 0, DN_ERR_INVALID, "bad GEMM shape");
    DN_REQUIRE(K % 8 == 0, DN_ERR_UNSUPPORTED, "GEMM K (input channels) must be a multiple of 8 (got %d)", K);
    DN_REQUIRE(impl == 0 || impl == 1, DN_ERR_INVALID, "impl must be 0 (tcgen05) or 1 (SIMT self-check)");
    PwEpilogue ep;
    ep.bias = bias;
    ep.residual = (const dn_half_t*)residual;
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
