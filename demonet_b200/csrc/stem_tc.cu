// Stem (input normalisation + dense 3x3 stride-2 conv + folded BN + activation) on the tcgen05 tensor cores, sm_100a.
// Replaces GeneralizedRCNNTransform.normalize (transform.py:129-138) followed by the first ConvBNActivation
// (mobilenetv3.py:141-142, mobilenetv2.py:157) -- same contract as stem_tma.cu, whose 27 x Cout fp32 FMAs per pixel made
// that kernel instruction-bound (0.28 of HBM peak).  Here the threads only BUILD the im2col rows; the multiply-adds run as
//
//   D[128 pixels, Cout] = A[128, 27 (padded to 32)] . W[Cout, 32]^T      (tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//
// with fp32 accuracy kept by a two-term split of BOTH operands into fp16 pairs, a = a_hi + a_lo, w = w_hi + w_lo
// (|x - x_hi - x_lo| <= 2^-22 |x|): D = A_hi W_hi + A_lo W_hi + A_hi W_lo, six 128 x Cout x 16 UMMAs per tile; the dropped
// A_lo W_lo term is below fp32 rounding.  The fp16 products are exact in the fp32 accumulator.
//
// A persistent CTA (128 threads) walks tiles of 4 x 32 output pixels.  Per tile:
//   TMA      one cp.async.bulk.tensor.3d fetches the 9 x 72 x 3 fp32 input window (NCHW planes) into a ring of four
//   im2col   thread = pixel: 27 taps are read from the window, normalised with the reference's exact (x - mean) / std
//            (the Markstein sequence of stem_tma.cu), zeroed where the convolution pads, split into hi / lo halves and
//            written as the K-major SWIZZLE_32B rows of four A operands (hi k0, hi k1, lo k0, lo k1)
//   MMA      one thread issues the six UMMAs; the accumulator of tile i is drained while tile i + 1 is being built
//   drain    tcgen05.ld (thread = pixel = TMEM lane) -> + bias -> activation -> 16-bit NHWC, 32 contiguous bytes per pixel
#include <cuda.h>

#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "dwconv.cuh"

namespace dn {

constexpr int SC_TH = 4, SC_TW = 32;                 // output tile = 128 pixels = one MMA tile
constexpr int SC_IH = 2 * SC_TH + 1;                 // 9 input rows
constexpr int SC_XOFF = 4;                           // window starts 4 columns left of the tile (16-byte aligned TMA start)
constexpr int SC_IW = 72;                            // 3 unused + 65 used input columns, padded to a 16-byte multiple
constexpr int SC_RAW_FLOATS = 3 * SC_IH * SC_IW;     // 1944 floats = 7776 bytes
constexpr int SC_RAW_BYTES = (SC_RAW_FLOATS * 4 + 127) & ~127;
constexpr int SC_A_TILE = 128 * 32;                  // one K = 16 operand: 128 rows of 32 bytes
constexpr int SC_A_BYTES = 4 * SC_A_TILE;            // hi k0, hi k1, lo k0, lo k1
constexpr int SC_THREADS = 128;
constexpr int SC_RING = 4;                           // input windows in flight per CTA: a window is needed one tile after its slot is
                                                     // freed, so two slots expose the whole DRAM latency (measured: 0.37 ms)

__device__ __forceinline__ uint32_t sc_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sc_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(sc_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
template <int ACT>
__device__ __forceinline__ float sc_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));      // x * relu6(x + 3) / 6
    return v;
}
// K-major SWIZZLE_32B operand descriptor, 8-row groups 256 bytes apart (the layout dwpw_fused.cu runs on)
__device__ __forceinline__ uint64_t sc_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
// fp32 pair -> packed fp16 pair (round to nearest), and back
__device__ __forceinline__ uint32_t sc_pack_f16(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 sc_unpack_f16(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

struct ScNorm {
    float mean[3], std[3], rinv[3];
};

struct __align__(8) ScBars {
    uint64_t full[SC_RING];    // TMA -> im2col: input window landed
    uint64_t mma_done[2];      // MMA -> drain: accumulator complete
    uint32_t tmem_base;
    uint32_t pad;
};

template <int COUT, int ACT>
__global__ void __launch_bounds__(SC_THREADS)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmap_img, const float* __restrict__ w, const float* __restrict__ bias,
               uint4* __restrict__ y, int H, int W, int Ho, int Wo, int tiles_x, int tiles_y, int n_tiles,
               const __grid_constant__ ScNorm nm) {
    constexpr int B_TILE = COUT * 32;                // one K = 16 weight operand: COUT rows of 32 bytes
    constexpr uint32_t TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;
    extern __shared__ __align__(1024) unsigned char sc_smem_raw[];
    unsigned char* base = sc_smem_raw + ((1024u - (sc_u32(sc_smem_raw) & 1023u)) & 1023u);
    unsigned char* a_t = base;                                       // [4][128 x 16] fp16, SWIZZLE_32B (single buffer: the six
                                                                     // UMMAs of a tile retire long before the next tile's rows are ready)
    unsigned char* w_t = a_t + SC_A_BYTES;                       // [4][COUT x 16] fp16: hi k0, hi k1, lo k0, lo k1
    float* raw = reinterpret_cast<float*>(w_t + 4 * B_TILE);         // [SC_RING][3][9][72] fp32 input windows
    float* sb = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(raw) + SC_RING * SC_RAW_BYTES);      // [COUT]
    ScBars* bars = reinterpret_cast<ScBars*>(sb + COUT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_img) : "memory");
        for (int i = 0; i < SC_RING; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sc_u32(&bars->full[i])) : "memory");
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sc_u32(&bars->mma_done[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sc_u32(&bars->tmem_base)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weights (not produced by a previous kernel): w[k][n] fp32, k = (ci * 3 + kh) * 3 + kw -> hi / lo halves of the B operands.
    // SWIZZLE_32B: the 16-byte chunk c of row n lives at chunk position c ^ ((n >> 2) & 1); rows k = 27..31 are zero.
    for (int i = threadIdx.x; i < COUT * 32; i += SC_THREADS) {
        const int n = i >> 5, k = i & 31;
        const float v = k < 27 ? __ldg(w + k * COUT + n) : 0.f;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const int kk = k & 15, c = kk >> 3;
        const int off = (k >> 4) * B_TILE + n * 32 + ((c ^ ((n >> 2) & 1)) << 4) + (kk & 7) * 2;
        *reinterpret_cast<__half*>(w_t + off) = hi;
        *reinterpret_cast<__half*>(w_t + 2 * B_TILE + off) = lo;
    }
    for (int i = threadIdx.x; i < COUT; i += SC_THREADS) sb[i] = __ldg(bias + i);
    pdl_trigger();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the weight tiles are read by the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    // Every CTA owns a CONTIGUOUS range of tiles (row-major over tx, ty, image): neighbouring tiles share their halo rows in
    // L2, and the tile coordinates advance by increments -- no integer division inside the loop (three divisions per tile
    // and thread were a third of the first version's instructions).
    struct Cursor {
        int tx, ty, b;
    };
    auto advance = [&](Cursor& q) {
        if (++q.tx == tiles_x) {
            q.tx = 0;
            if (++q.ty == tiles_y) q.ty = 0, ++q.b;
        }
    };
    const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x), t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    Cursor cur;
    cur.tx = t0 % tiles_x, cur.ty = (t0 / tiles_x) % tiles_y, cur.b = t0 / (tiles_x * tiles_y);

    auto issue_tile = [&](const Cursor& q, int slot) {
        const uint32_t bar = sc_u32(&bars->full[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SC_RAW_FLOATS * 4)) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                sc_u32(reinterpret_cast<unsigned char*>(raw) + slot * SC_RAW_BYTES)),
            "l"(&tmap_img), "r"(bar), "r"(q.tx * SC_TW * 2 - SC_XOFF), "r"(q.ty * SC_TH * 2 - 1), "r"(q.b * 3)
            : "memory");
    };
    Cursor pre = cur;                                // thread 0: the window SC_RING tiles ahead
    int t_pre = t0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < SC_RING && t_pre < t1; ++i, ++t_pre) {
            issue_tile(pre, i);
            advance(pre);
        }
    }

    // drain of one finished tile: thread = pixel (TMEM lane 32 * warp + lane = row `warp` of the tile, column `lane`)
    auto drain = [&](const Cursor& q, uint32_t dit) {
        const uint32_t dbuf = dit & 1u;
        sc_wait(&bars->mma_done[dbuf], (dit >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int oh = q.ty * SC_TH + warp, ow = q.tx * SC_TW + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + dbuf * (uint32_t)COUT;
        uint32_t v[COUT];
#pragma unroll
        for (int j = 0; j < COUT / 8; ++j)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[j * 8 + 0]), "=r"(v[j * 8 + 1]), "=r"(v[j * 8 + 2]), "=r"(v[j * 8 + 3]), "=r"(v[j * 8 + 4]),
                           "=r"(v[j * 8 + 5]), "=r"(v[j * 8 + 6]), "=r"(v[j * 8 + 7])
                         : "r"(taddr + (uint32_t)(j * 8)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (oh < Ho && ow < Wo) {
            uint4* yo = y + (((long long)q.b * Ho + oh) * Wo + ow) * (COUT / 8);
#pragma unroll
            for (int j = 0; j < COUT / 8; ++j) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = sc_act<ACT>(__uint_as_float(v[j * 8 + e]) + sb[j * 8 + e]);
                yo[j] = pack8(f);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    const int r = warp, c = lane;                    // this thread's pixel inside a tile
    const int m = threadIdx.x;                       // its row in the A operands
    const uint32_t a_row = (uint32_t)(m * 32);
    const uint32_t sw = (uint32_t)((m >> 2) & 1);
    // im2col of this thread's pixel: 27 normalised taps, K order (ci, kh, kw).  CHECK = false for tiles whose windows lie
    // inside the image (no tap is padding: the comparisons and selects go away).
    auto gather = [&](const float* win, int ih0, int iw0, auto check, float* a) {
        constexpr bool CHECK = decltype(check)::value;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            const float mu = nm.mean[ci], sd = nm.std[ci], ri = nm.rinv[ci];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const bool row_ok = !CHECK || (unsigned)(ih0 + kh) < (unsigned)H;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float x = win[(ci * SC_IH + kh) * SC_IW + kw];
                    const float d = __fsub_rn(x, mu);
                    const float q0 = __fmul_rn(d, ri);
                    const float e = __fmaf_rn(-q0, sd, d);
                    const float q = __fmaf_rn(e, ri, q0);                         // = RN(d / std)
                    a[(ci * 3 + kh) * 3 + kw] = (!CHECK || (row_ok && (unsigned)(iw0 + kw) < (unsigned)W)) ? q : 0.f;
                }
            }
        }
    };
    uint32_t it = 0;
    Cursor prev = cur;
    for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t buf = it & 1u, slot = it % SC_RING;
        const int ih0 = cur.ty * SC_TH * 2 - 1 + 2 * r, iw0 = (cur.tx * SC_TW + c) * 2 - 1;
        // CTA-uniform: does the 9 x 65 window of this tile lie inside the image?
        const bool interior = cur.ty > 0 && cur.tx > 0 && (cur.ty * SC_TH * 2 - 1 + SC_IH <= H) && (cur.tx * SC_TW * 2 - 1 + 2 * SC_TW + 1 <= W);
        sc_wait(&bars->full[slot], (it / SC_RING) & 1u);
        const float* win = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(raw) + slot * SC_RAW_BYTES) +
                           (2 * r) * SC_IW + (SC_XOFF - 1) + 2 * c;
        float a[32];
        if (interior) gather(win, ih0, iw0, std::false_type{}, a);
        else gather(win, ih0, iw0, std::true_type{}, a);
#pragma unroll
        for (int k = 27; k < 32; ++k) a[k] = 0.f;
        unsigned char* at = a_t;
        // the previous tile's UMMAs read the A operands: they must have retired (they were issued a whole drain + gather ago)
        if (it > 0) sc_wait(&bars->mma_done[(it - 1) & 1u], ((it - 1) >> 1) & 1u);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {             // K step (16 taps) -> one hi and one lo operand
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {         // 16-byte chunk = 8 taps
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float x0 = a[ks * 16 + ch * 8 + 2 * p], x1 = a[ks * 16 + ch * 8 + 2 * p + 1];
                    hi[p] = sc_pack_f16(x0, x1);
                    const float2 back = sc_unpack_f16(hi[p]);
                    lo[p] = sc_pack_f16(x0 - back.x, x1 - back.y);
                }
                const uint32_t off = a_row + (((uint32_t)ch ^ sw) << 4);
                *reinterpret_cast<uint4*>(at + ks * SC_A_TILE + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(at + (2 + ks) * SC_A_TILE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the MMA
        __syncthreads();                             // A operands complete; the window buffer is free; TMEM[buf] was drained
        if (threadIdx.x == 0) {
            if (t_pre < t1) {
                issue_tile(pre, (int)slot);
                advance(pre);
                ++t_pre;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // instruction descriptor: D = f32, A = B = f16 (format 0), K-major both, N = COUT, M = 128
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t d = tmem_base + buf * (uint32_t)COUT;
            const uint32_t a0 = sc_u32(at), b0 = sc_u32(w_t);
            // (A operand, B operand): hi0 x hi0, hi1 x hi1, lo0 x hi0, lo1 x hi1, hi0 x lo0, hi1 x lo1
            const int ai[6] = {0, 1, 2, 3, 0, 1}, bi[6] = {0, 1, 0, 1, 2, 3};
#pragma unroll
            for (int j = 0; j < 6; ++j)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                    "l"(sc_desc_sw32(a0 + ai[j] * SC_A_TILE)), "l"(sc_desc_sw32(b0 + bi[j] * B_TILE)), "r"(idesc), "r"(j ? 1u : 0u)
                    : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sc_u32(&bars->mma_done[buf]))
                         : "memory");
        }
        // the previous tile's accumulator finished long ago: drain it while this tile's MMAs run
        if (it > 0) drain(prev, it - 1);
        prev = cur;
        advance(cur);
    }
    if (it > 0) drain(prev, it - 1);
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledSc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledSc sc_encode_fn() {
    static PFN_encodeTiledSc fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiledSc)p;
    }
    return fn;
}

static int stem_tc_make_tmap(CUtensorMap* map, const float* images, int B, int H, int W) {
    PFN_encodeTiledSc fn = sc_encode_fn();
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
    cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {(cuuint32_t)SC_IW, (cuuint32_t)SC_IH, 3};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(images), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled (tensor-core stem) failed (%d): B=%d H=%d W=%d", (int)r, B, H, W);
    return DN_OK;
}

template <int COUT, int ACT>
static int stem_tc_launch_t(const CUtensorMap& tm, const float* w, const float* bias, const ScNorm& nm, void* y, int B, int H, int W,
                            cudaStream_t stream) {
    auto kern = stem_tc_kernel<COUT, ACT>;
    const size_t smem = SC_A_BYTES + 4 * COUT * 32 + SC_RING * SC_RAW_BYTES + COUT * 4 + sizeof(ScBars) + 1024;
    static SmemOptIn optin;
    int per_sm = 1;
    DN_CHECK_CUDA(optin.ensure(kern, smem));
    DN_CHECK_CUDA(optin.blocks_per_sm_tmem(kern, SC_THREADS, smem, 2 * COUT < 32 ? 32 : 2 * COUT, &per_sm));
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int tiles_x = ceil_div(Wo, SC_TW), tiles_y = ceil_div(Ho, SC_TH);
    const long long n_tiles = (long long)B * tiles_x * tiles_y;
    DN_REQUIRE(n_tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "stem problem too large");
    long long grid = (long long)per_sm * sm_count();
    if (grid > n_tiles) grid = n_tiles;
    launch_pdl(kern, (unsigned)grid, SC_THREADS, smem, stream, tm, w, bias, (uint4*)y, H, W, Ho, Wo, tiles_x, tiles_y, (int)n_tiles, nm);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

template <int COUT>
static int stem_tc_launch_a(const CUtensorMap& tm, const float* w, const float* bias, const ScNorm& nm, void* y, int B, int H, int W,
                            int act, cudaStream_t stream) {
    switch (act) {
        case DN_ACT_RELU: return stem_tc_launch_t<COUT, DN_ACT_RELU>(tm, w, bias, nm, y, B, H, W, stream);
        case DN_ACT_RELU6: return stem_tc_launch_t<COUT, DN_ACT_RELU6>(tm, w, bias, nm, y, B, H, W, stream);
        case DN_ACT_HSWISH: return stem_tc_launch_t<COUT, DN_ACT_HSWISH>(tm, w, bias, nm, y, B, H, W, stream);
        default: return stem_tc_launch_t<COUT, DN_ACT_NONE>(tm, w, bias, nm, y, B, H, W, stream);
    }
}

// DN_STEM=simt selects the fp32 SIMT kernel of stem_tma.cu (A/B measurements); default: tensor cores
bool stem_tc_enabled() {
    const char* e = getenv("DN_STEM");               // read per call: the tests flip it inside one process
    return !(e && strcmp(e, "simt") == 0);
}

int stem_tc_launch(const float* images, const float* w, const float* bias, const float* mean3, const float* std3, void* y, int B,
                   int H, int W, int Cout, int act, cudaStream_t stream) {
    ScNorm nm;
    for (int i = 0; i < 3; ++i) {
        nm.mean[i] = mean3[i], nm.std[i] = std3[i];
        nm.rinv[i] = (float)(1.0 / (double)std3[i]);                   // stem_norm_ok() has vetted the divisors
    }
    CUtensorMap tm;
    int rc = stem_tc_make_tmap(&tm, images, B, H, W);
    if (rc) return rc;
    if (Cout == 16) return stem_tc_launch_a<16>(tm, w, bias, nm, y, B, H, W, act, stream);
    return stem_tc_launch_a<32>(tm, w, bias, nm, y, B, H, W, act, stream);
}

}  // namespace dn
