// Depthwise k x k convolution as a TMA-fed, shared-memory-tiled stencil for sm_100a.
//
// A persistent CTA walks (image, tile_y, tile_x) output tiles of one channel chunk.  The input halo tile
// (IHT x IWT pixels x CB channels, NHWC bf16) is fetched by ONE cp.async.bulk.tensor.4d per tile into a
// double-buffered shared-memory slot -- the zero padding of the convolution is the tensor map's
// out-of-bounds fill, so there are no boundary branches -- and tile i+1 is in flight while tile i is
// consumed.  Thread = 8 channels x TW output columns of one output row; the stencil reads shared memory
// (one 16-byte load per input vector, unpacked once, reused by up to k taps x TW outputs from registers),
// accumulates in fp32 with packed FFMA2, applies folded-BN bias + activation and stores bf16.
// Reference: ConvBNActivation(groups=C), demonet/models/mobilenetv2.py:32-55 (see dn_dwconv).
#include <cuda.h>

#include "common.cuh"
#include "dwconv.cuh"

namespace dn {

__device__ __forceinline__ uint32_t dsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ACT>
__device__ __forceinline__ float dwt_act(float v) {
    if (ACT == DN_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == DN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (ACT == DN_ACT_HSWISH) return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f));      // x * relu6(x + 3) / 6
    return v;
}

template <int KS, int S, int ACT>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ w, const float* __restrict__ bias,
                  uint4* __restrict__ y, DwTiling tl, int C, int Ho, int Wo, int n_tiles) {
    constexpr int PAD = (KS - 1) / 2;
    constexpr int TW = (S == 1) ? 4 : 2;
    constexpr int NV = (TW - 1) * S + KS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tile_bytes = tl.IHT * tl.IWT * tl.CB * 2;
    const int tile_stride = (tile_bytes + 127) & ~127;
    unsigned char* tiles = smem;
    float* wsm = reinterpret_cast<float*>(smem + 2 * tile_stride);      // [KS*KS][CB]
    float* bsm = wsm + KS * KS * tl.CB;                                  // [CB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + tl.CB);           // [2]

    const int chunk = blockIdx.y;
    const int c0 = chunk * tl.CB;
    const int cvs = tl.CB >> 3;
    const int CV = C >> 3;

    auto issue = [&](int t, int buf) {
        const int tx = t % tl.tiles_x;
        const int ty = (t / tl.tiles_x) % tl.tiles_y;
        const int b = t / (tl.tiles_x * tl.tiles_y);
        const uint32_t bar = dsmem_u32(&bars[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)tile_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                "r"(dsmem_u32(tiles + buf * tile_stride)),
            "l"(&tmap_x), "r"(bar), "r"(c0), "r"(tx * tl.TWo * S - PAD), "r"(ty * tl.THo * S - PAD), "r"(b)
            : "memory");
    };

    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dsmem_u32(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dsmem_u32(&bars[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    }
    for (int i = threadIdx.x; i < KS * KS * tl.CB; i += DW_THREADS) wsm[i] = w[(i / tl.CB) * C + c0 + (i % tl.CB)];
    for (int i = threadIdx.x; i < tl.CB; i += DW_THREADS) bsm[i] = bias[c0 + i];
    __syncthreads();

    const int cv = threadIdx.x % cvs;
    const int strip = threadIdx.x / cvs;
    const bool active = strip < tl.THo * tl.spr;
    const int r = strip / tl.spr;                       // local output row
    const int ol0 = (strip % tl.spr) * TW;              // local output column of the strip
    const float4* wv4 = reinterpret_cast<const float4*>(wsm) + cv * 2;
    const int w_pitch4 = tl.CB >> 2;                    // float4 per tap row

    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (threadIdx.x == 0 && t + (int)gridDim.x < n_tiles) issue(t + gridDim.x, buf ^ 1);
        {   // wait for this tile's bytes
            const uint32_t bar = dsmem_u32(&bars[buf]);
            const uint32_t parity = (it >> 1) & 1u;
            uint32_t ok = 0;
            do {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(bar), "r"(parity)
                    : "memory");
            } while (!ok);
        }
        if (active) {
            const int tx = t % tl.tiles_x;
            const int ty = (t / tl.tiles_x) % tl.tiles_y;
            const int b = t / (tl.tiles_x * tl.tiles_y);
            const uint4* tin = reinterpret_cast<const uint4*>(tiles + buf * tile_stride);
            float2 acc[TW][4];
            {
                const float4 b0 = reinterpret_cast<const float4*>(bsm)[cv * 2], b1 = reinterpret_cast<const float4*>(bsm)[cv * 2 + 1];
#pragma unroll
                for (int q = 0; q < TW; ++q) {
                    acc[q][0] = make_float2(b0.x, b0.y); acc[q][1] = make_float2(b0.z, b0.w);
                    acc[q][2] = make_float2(b1.x, b1.y); acc[q][3] = make_float2(b1.z, b1.w);
                }
            }
#pragma unroll
            for (int kh = 0; kh < KS; ++kh) {
                const uint4* row = tin + ((r * S + kh) * tl.IWT + ol0 * S) * cvs + cv;
                float2 in[NV][4];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const uint4 v = row[i * cvs];
                    in[i][0] = h2_to_float2(v.x);
                    in[i][1] = h2_to_float2(v.y);
                    in[i][2] = h2_to_float2(v.z);
                    in[i][3] = h2_to_float2(v.w);
                }
#pragma unroll
                for (int kw = 0; kw < KS; ++kw) {
                    const float4 w0 = wv4[(kh * KS + kw) * w_pitch4], w1 = wv4[(kh * KS + kw) * w_pitch4 + 1];
                    const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                          make_float2(w1.z, w1.w)};
#pragma unroll
                    for (int q = 0; q < TW; ++q) {
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc[q][p] = __ffma2_rn(in[q * S + kw][p], wv[p], acc[q][p]);
                    }
                }
            }
            const int oh = ty * tl.THo + r;
            const int ow0 = tx * tl.TWo + ol0;
            if (oh < Ho) {
                uint4* yo = y + (((long long)b * Ho + oh) * Wo + ow0) * CV + chunk * cvs + cv;
#pragma unroll
                for (int q = 0; q < TW; ++q) {
                    if (ow0 + q < Wo) {
                        uint4 o;
                        o.x = float2_to_h2(dwt_act<ACT>(acc[q][0].x), dwt_act<ACT>(acc[q][0].y));
                        o.y = float2_to_h2(dwt_act<ACT>(acc[q][1].x), dwt_act<ACT>(acc[q][1].y));
                        o.z = float2_to_h2(dwt_act<ACT>(acc[q][2].x), dwt_act<ACT>(acc[q][2].y));
                        o.w = float2_to_h2(dwt_act<ACT>(acc[q][3].x), dwt_act<ACT>(acc[q][3].y));
                        yo[(long long)q * CV] = o;
                    }
                }
            }
        }
        __syncthreads();        // everyone is done with tiles[buf] before it is refilled two iterations later
    }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
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

int dw_plan(int H, int W, int C, int k, int stride, DwTiling* tl) {
    const int pad = (k - 1) / 2;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    const int TW = stride == 1 ? 4 : 2;
    int CB = 8;
    for (int c = 64; c >= 8; c -= 8)
        if (C % c == 0) {
            CB = c;
            break;
        }
    const int cvs = CB / 8;
    const int strips_max = DW_THREADS / cvs;
    const int strips_w = (Wo + TW - 1) / TW;
    // pick (strips per row, rows) by brute force: minimise the worse of the two inefficiencies -- compute
    // slots per useful output strip and halo-tile pixels loaded per input pixel -- within the smem budget
    auto tile_bytes = [&](int th, int sp) { return ((th - 1) * stride + k) * ((sp * TW - 1) * stride + k) * CB * 2; };
    int best_spr = 1, best_tho = 1;
    double best_cost = 1e30;
    for (int sp = 1; sp <= strips_w && sp <= strips_max; ++sp) {
        int th = strips_max / sp;
        if (th > Ho) th = Ho;
        if (th > 16) th = 16;
        while (th > 1 && 2 * tile_bytes(th, sp) > 96 * 1024) --th;
        if (((sp * TW - 1) * stride + k) > 256) continue;
        const int tx = (strips_w + sp - 1) / sp, ty = (Ho + th - 1) / th;
        const double slots = (double)tx * ty * (strips_max) / ((double)Ho * strips_w);      // thread slots per useful strip
        const double loads = (double)tx * ty * ((th - 1) * stride + k) * ((sp * TW - 1) * stride + k) / ((double)H * W);
        const double cost = (slots > loads ? slots : loads) + 0.05 * (slots < loads ? slots : loads);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best_spr = sp;
            best_tho = th;
        }
    }
    const int spr = best_spr;
    int tho = best_tho;
    tl->CB = CB;
    tl->THo = tho;
    tl->spr = spr;
    tl->TWo = spr * TW;
    tl->IHT = (tho - 1) * stride + k;
    tl->IWT = (tl->TWo - 1) * stride + k;
    tl->tiles_x = (Wo + tl->TWo - 1) / tl->TWo;
    tl->tiles_y = (Ho + tho - 1) / tho;
    tl->chunks = C / CB;
    DN_REQUIRE(tl->IWT <= 256 && tl->IHT <= 256, DN_ERR_UNSUPPORTED, "depthwise tile too large for a TMA box");
    DN_REQUIRE(2 * tile_bytes(tho, spr) <= 200 * 1024, DN_ERR_UNSUPPORTED, "depthwise tile too large for shared memory");
    (void)pad;
    return DN_OK;
}

int dw_make_tmap(CUtensorMap* map, const void* x, int B, int H, int W, int C, const DwTiling& tl) {
    PFN_encodeTiled fn = encode_fn();
    DN_REQUIRE(fn != nullptr, DN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    DN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, DN_ERR_INVALID, "depthwise input must be 16-byte aligned");
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)tl.CB, (cuuint32_t)tl.IWT, (cuuint32_t)tl.IHT, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, DN_TMAP_HALF, 4, const_cast<void*>(x), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DN_REQUIRE(r == CUDA_SUCCESS, DN_ERR_CUDA, "cuTensorMapEncodeTiled (4D) failed (%d): C=%d W=%d H=%d B=%d box=%d,%d,%d",
               (int)r, C, W, H, B, tl.CB, tl.IWT, tl.IHT);
    return DN_OK;
}

template <int KS, int S, int ACT>
static int launch(const CUtensorMap& tm, const float* w, const float* bias, void* y, const DwTiling& tl, int B, int C, int Ho,
                  int Wo, cudaStream_t stream) {
    const int tile_bytes = tl.IHT * tl.IWT * tl.CB * 2;
    const size_t smem = 2 * (size_t)((tile_bytes + 127) & ~127) + (size_t)(KS * KS + 1) * tl.CB * 4 + 16;
    static SmemOptIn optin;
    DN_CHECK_CUDA(optin.ensure(dwconv_tma_kernel<KS, S, ACT>, smem));
    const long long n_tiles = (long long)B * tl.tiles_x * tl.tiles_y;
    DN_REQUIRE(n_tiles < (1ll << 31), DN_ERR_UNSUPPORTED, "depthwise problem too large");
    // persistent: about two CTAs per SM in total, spread over the channel chunks
    int gx = (2 * sm_count() + tl.chunks - 1) / tl.chunks;
    if (gx > n_tiles) gx = (int)n_tiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, tl.chunks);
    launch_pdl(dwconv_tma_kernel<KS, S, ACT>, grid, DW_THREADS, smem, stream, tm, w, bias, (uint4*)y, tl, C, Ho, Wo, (int)n_tiles);
    DN_CHECK_LAUNCH();
    return DN_OK;
}

template <int KS, int S>
static int launch_act(const CUtensorMap& tm, const float* w, const float* bias, void* y, const DwTiling& tl, int B, int C,
                      int Ho, int Wo, int act, cudaStream_t stream) {
    switch (act) {
        case DN_ACT_RELU: return launch<KS, S, DN_ACT_RELU>(tm, w, bias, y, tl, B, C, Ho, Wo, stream);
        case DN_ACT_RELU6: return launch<KS, S, DN_ACT_RELU6>(tm, w, bias, y, tl, B, C, Ho, Wo, stream);
        case DN_ACT_HSWISH: return launch<KS, S, DN_ACT_HSWISH>(tm, w, bias, y, tl, B, C, Ho, Wo, stream);
        default: return launch<KS, S, DN_ACT_NONE>(tm, w, bias, y, tl, B, C, Ho, Wo, stream);
    }
}

int dwconv_tma_launch(const CUtensorMap& tm, const DwTiling& tl, const float* w, const float* bias, void* y, int B, int H,
                      int W, int C, int k, int stride, int act, cudaStream_t stream) {
    const int pad = (k - 1) / 2;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    if (k == 3 && stride == 1) return launch_act<3, 1>(tm, w, bias, y, tl, B, C, Ho, Wo, act, stream);
    if (k == 3 && stride == 2) return launch_act<3, 2>(tm, w, bias, y, tl, B, C, Ho, Wo, act, stream);
    if (k == 5 && stride == 1) return launch_act<5, 1>(tm, w, bias, y, tl, B, C, Ho, Wo, act, stream);
    return launch_act<5, 2>(tm, w, bias, y, tl, B, C, Ho, Wo, act, stream);
}

}  // namespace dn
