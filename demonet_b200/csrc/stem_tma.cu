// Stem (input normalisation + dense 3x3 stride-2 conv + folded BN + activation) as a TMA-fed
// shared-memory-tiled kernel for sm_100a.  Replaces GeneralizedRCNNTransform.normalize (transform.py:129-138)
// followed by the first ConvBNActivation (mobilenetv3.py:141-142, mobilenetv2.py:157).
//
// A persistent CTA walks 16 x 32 output tiles.  ONE cp.async.bulk.tensor.3d per tile fetches the 33 x 72 x 3
// fp32 input window (NCHW planes) into `raw`; the window is normalised ONCE per element into `nrm`
// (out-of-image elements become the conv's zero padding), `raw` is then free and the next tile's TMA is
// issued, overlapping the stencil.  Thread = NPX adjacent output pixels x all COUT channels (NPX * COUT = 64
// fp32 accumulators): taps are vector loads from `nrm`, weights are broadcast LDS.128 shared by the NPX
// pixels, the math is packed FFMA2.
//
// (x - mean) / std is computed WITHOUT the 12-instruction division sequence but with the same bits: with
// r = RN(1/std) precomputed on the host, q0 = RN(d*r), e = fma(-q0, std, d) (exact), q = fma(e, r, q0) is the
// correctly rounded quotient d/std (Markstein's theorem; needs a correctly rounded reciprocal and a divisor
// whose significand is not all ones -- checked on the host, which otherwise uses the direct kernel).
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "dwconv.cuh"

namespace dn {

constexpr int ST_TH = 16, ST_TW = 32;                // output tile
constexpr int ST_IH = 2 * ST_TH + 1;                 // 33 input rows
constexpr int ST_XOFF = 4;                           // the window starts 4 columns left of the tile, not 1: TMA needs the
                                                     // innermost start coordinate 16-byte aligned (measured: x = 63 or -1
                                                     // raises 'illegal instruction', x = 64 or -4 is fine)
constexpr int ST_IW = 72;                            // 3 unused + 65 used input columns, padded to a 16-byte multiple
constexpr int ST_NW = 68;                            // normalised row: the 65 used columns (+3 pad), 16-byte aligned rows
constexpr int ST_TILE_FLOATS = 3 * ST_IH * ST_IW;
constexpr int ST_RAW_FLOATS = (ST_TILE_FLOATS + 31) / 32 * 32;        // TMA destinations must be 128-byte aligned
constexpr int ST_NRM_FLOATS = 3 * ST_IH * ST_NW;

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct StemNorm {
    float mean[3], std[3], rinv[3];
};

template <int ACT>
__device__ __forceinline__ float st_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));      // x * relu6(x + 3) / 6
    return v;
}

template <int COUT, int NPX, int ACT>
__global__ void __launch_bounds__(ST_TH * (ST_TW / NPX))
stem_tma_kernel(const __grid_constant__ CUtensorMap tmap_img, const float* __restrict__ w, const float* __restrict__ bias,
                uint4* __restrict__ y, int H, int W, int Ho, int Wo, int tiles_x, int tiles_y, int n_tiles,
                const __grid_constant__ StemNorm nm) {
    constexpr int QW = ST_TW / NPX;                  // threads across a tile row
    constexpr int THREADS = ST_TH * QW;
    constexpr int NV = 2 * NPX + 1;                  // input columns one thread touches per (ci, kh)
    extern __shared__ __align__(128) float st_smem[];
    float* raw = st_smem;
    float* nrm = raw + ST_RAW_FLOATS;
    float* sw = nrm + ST_NRM_FLOATS;
    float* sb = sw + 27 * COUT;
    __shared__ uint64_t bar;

    auto issue = [&](int t) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const uint32_t ba = st_smem_u32(&bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"((uint32_t)(ST_TILE_FLOATS * 4))
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                st_smem_u32(raw)),
            "l"(&tmap_img), "r"(ba), "r"(tx * ST_TW * 2 - ST_XOFF), "r"(ty * ST_TH * 2 - 1), "r"(b * 3)
            : "memory");
    };
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int)blockIdx.x < n_tiles) issue(blockIdx.x);
    }
    for (int i = threadIdx.x; i < 27 * COUT; i += THREADS) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += THREADS) sb[i] = bias[i];
    __syncthreads();

    const int r = threadIdx.x / QW, q = threadIdx.x % QW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        {
            const uint32_t ba = st_smem_u32(&bar), parity = it & 1u;
            uint32_t ok = 0;
            do {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(ba), "r"(parity)
                    : "memory");
            } while (!ok);
        }
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const int ih0 = ty * ST_TH * 2 - 1, iw0 = tx * ST_TW * 2 - 1;
        // normalise raw -> nrm: a warp per (plane, row), lanes across the columns; everything but the column test is
        // warp-uniform.  nrm column j holds image column iw0 + j.
        for (int row = warp; row < 3 * ST_IH; row += THREADS / 32) {
            const int ci = row / ST_IH, ih = ih0 + (row - ci * ST_IH);
            const bool row_ok = (ih >= 0) && (ih < H);
            const float m = ci == 0 ? nm.mean[0] : (ci == 1 ? nm.mean[1] : nm.mean[2]);
            const float sd = ci == 0 ? nm.std[0] : (ci == 1 ? nm.std[1] : nm.std[2]);
            const float ri = ci == 0 ? nm.rinv[0] : (ci == 1 ? nm.rinv[1] : nm.rinv[2]);
            const float* src = raw + row * ST_IW + (ST_XOFF - 1);
            float* dst = nrm + row * ST_NW;
#pragma unroll
            for (int j0 = 0; j0 < ST_NW; j0 += 32) {
                const int j = j0 + lane;
                if (j < ST_NW) {
                    const int iw = iw0 + j;
                    const float d = __fsub_rn(src[j], m);
                    const float q0 = __fmul_rn(d, ri);
                    const float e = __fmaf_rn(-q0, sd, d);
                    const float qq = __fmaf_rn(e, ri, q0);
                    dst[j] = (row_ok && iw >= 0 && iw < W) ? qq : 0.f;
                }
            }
        }
        __syncthreads();                                     // nrm complete; raw is free again
        if (threadIdx.x == 0 && t + (int)gridDim.x < n_tiles) issue(t + gridDim.x);

        float2 acc[NPX][COUT / 2];
#pragma unroll
        for (int p = 0; p < NPX; ++p)
#pragma unroll
            for (int c2 = 0; c2 < COUT / 2; ++c2) acc[p][c2] = make_float2(sb[2 * c2], sb[2 * c2 + 1]);
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* rowp = nrm + (ci * ST_IH + 2 * r + kh) * ST_NW + 2 * NPX * q;
                float v[NV];
#pragma unroll
                for (int j4 = 0; j4 < NV / 4; ++j4) {
                    const float4 f = *reinterpret_cast<const float4*>(rowp + 4 * j4);
                    v[4 * j4 + 0] = f.x, v[4 * j4 + 1] = f.y, v[4 * j4 + 2] = f.z, v[4 * j4 + 3] = f.w;
                }
                v[NV - 1] = rowp[NV - 1];
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4* wr = reinterpret_cast<const float4*>(sw + ((ci * 3 + kh) * 3 + kw) * COUT);
#pragma unroll
                    for (int c4 = 0; c4 < COUT / 4; ++c4) {
                        const float4 ww = wr[c4];
#pragma unroll
                        for (int p = 0; p < NPX; ++p) {
                            const float x = v[2 * p + kw];
                            acc[p][c4 * 2 + 0] = __ffma2_rn(make_float2(x, x), make_float2(ww.x, ww.y), acc[p][c4 * 2 + 0]);
                            acc[p][c4 * 2 + 1] = __ffma2_rn(make_float2(x, x), make_float2(ww.z, ww.w), acc[p][c4 * 2 + 1]);
                        }
                    }
                }
            }
        }
        const int oh = ty * ST_TH + r, ow0 = tx * ST_TW + q * NPX;
        if (oh < Ho) {
            uint4* yo = y + (((long long)b * Ho + oh) * Wo + ow0) * (COUT / 8);
#pragma unroll
            for (int p = 0; p < NPX; ++p) {
                if (ow0 + p < Wo) {
#pragma unroll
                    for (int v8 = 0; v8 < COUT / 8; ++v8) {
                        uint4 o;
                        o.x = float2_to_h2(st_act<ACT>(acc[p][v8 * 4 + 0].x), st_act<ACT>(acc[p][v8 * 4 + 0].y));
                        o.y = float2_to_h2(st_act<ACT>(acc[p][v8 * 4 + 1].x), st_act<ACT>(acc[p][v8 * 4 + 1].y));
                        o.z = float2_to_h2(st_act<ACT>(acc[p][v8 * 4 + 2].x), st_act<ACT>(acc[p][v8 * 4 + 2].y));
                        o.w = float2_to_h2(st_act<ACT>(acc[p][v8 * 4 + 3].x), st_act<ACT>(acc[p][v8 * 4 + 3].y));
                        yo[p * (COUT / 8) + v8] = o;
                    }
                }
            }
        }
        __syncthreads();                                     // nrm is rewritten by the next tile's normalisation
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled st_encode_fn() {
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

// the TMA path needs 16-byte aligned image rows (W % 4 == 0) and a 16-byte aligned base
bool stem_can_tma(const void* images, int W) { return (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(images) & 15) == 0); }

int stem_make_tmap(CUtensorMap* map, const float* images, int B, int H, int W) {
    PFN_encodeTiled fn = st_encode_fn();
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
    cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {(cuuint32_t)ST_IW, (cuuint32_t)ST_IH, 3};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(images), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled (stem) failed (%d): B=%d H=%d W=%d", (int)r, B, H, W);
    return DN_OK;
}

template <int COUT, int NPX, int ACT>
static int stem_launch_t(const CUtensorMap& tm, const float* w, const float* bias, const StemNorm& nm, void* y, int B, int H,
                         int W, cudaStream_t stream) {
    constexpr int THREADS = ST_TH * (ST_TW / NPX);
    const size_t smem = (size_t)(ST_RAW_FLOATS + ST_NRM_FLOATS + 28 * COUT) * 4;
    static SmemOptIn optin;
    int ctas_per_sm = 1;
    DN_CHECK_CUDA(optin.ensure(stem_tma_kernel<COUT, NPX, ACT>, smem));
    DN_CHECK_CUDA(optin.blocks_per_sm(stem_tma_kernel<COUT, NPX, ACT>, THREADS, smem, &ctas_per_sm));
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int tiles_x = (Wo + ST_TW - 1) / ST_TW, tiles_y = (Ho + ST_TH - 1) / ST_TH;
    const long long n_tiles = (long long)B * tiles_x * tiles_y;
    DN_REQUIRE(n_tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "stem problem too large");
    long long grid = (long long)ctas_per_sm * sm_count();
    if (grid > n_tiles) grid = n_tiles;
    launch_pdl(stem_tma_kernel<COUT, NPX, ACT>, (unsigned)grid, THREADS, smem, stream, tm, w, bias, (uint4*)y, H, W, Ho, Wo, tiles_x,
               tiles_y, (int)n_tiles, nm);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

template <int COUT, int NPX>
static int stem_launch_a(const CUtensorMap& tm, const float* w, const float* bias, const StemNorm& nm, void* y, int B, int H,
                         int W, int act, cudaStream_t stream) {
    switch (act) {
        case DN_ACT_RELU: return stem_launch_t<COUT, NPX, DN_ACT_RELU>(tm, w, bias, nm, y, B, H, W, stream);
        case DN_ACT_RELU6: return stem_launch_t<COUT, NPX, DN_ACT_RELU6>(tm, w, bias, nm, y, B, H, W, stream);
        case DN_ACT_HSWISH: return stem_launch_t<COUT, NPX, DN_ACT_HSWISH>(tm, w, bias, nm, y, B, H, W, stream);
        default: return stem_launch_t<COUT, NPX, DN_ACT_NONE>(tm, w, bias, nm, y, B, H, W, stream);
    }
}

// r = RN(1/s); false when the shortcut division is not provably exact for this divisor
static bool stem_recip(float s, float* r) {
    uint32_t bits;
    memcpy(&bits, &s, 4);
    const uint32_t expo = (bits >> 23) & 0xff, mant = bits & 0x7fffff;
    if (!(s > 0.f) || expo == 0 || expo == 0xff || mant == 0x7fffff) return false;
    *r = (float)(1.0 / (double)s);
    return true;
}

bool stem_norm_ok(const float* std3) {
    float r;
    return stem_recip(std3[0], &r) && stem_recip(std3[1], &r) && stem_recip(std3[2], &r);
}

int stem_tma_launch(const CUtensorMap& tm, const float* w, const float* bias, const float* mean3, const float* std3, void* y,
                    int B, int H, int W, int Cout, int act, cudaStream_t stream) {
    StemNorm nm;
    for (int i = 0; i < 3; ++i) {
        nm.mean[i] = mean3[i], nm.std[i] = std3[i];
        DN_REQUIRE(stem_recip(std3[i], &nm.rinv[i]), DN_ERR_INVALID, "image_std[%d] = %g is not usable by the tiled stem", i,
                   (double)std3[i]);
    }
    if (Cout == 16) return stem_launch_a<16, 4>(tm, w, bias, nm, y, B, H, W, act, stream);
    return stem_launch_a<32, 2>(tm, w, bias, nm, y, B, H, W, act, stream);
}

}  // namespace dn
